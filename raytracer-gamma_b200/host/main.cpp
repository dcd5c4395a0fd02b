/* main.cpp — the host program of raytracer-gamma with the OpenCL plumbing
 * (reference main.cpp:182-350, 456-489, device_info.cpp, device_picker.h) replaced by
 * the C-ABI of include/rt_cuda.h.  Call sequence and output are the reference's:
 * build the scene (main.cpp:105-168), render, print "Exec time", read the float
 * pixels back, normalise by the global maximum (algebra.h:68-91) and write a
 * binary PPM (main.cpp:43-91).  Unlike the reference it takes its constants from
 * the command line and does not wait for a key press (main.cpp:505).
 *
 *   rt_gamma [--width W] [--height H] [--alias A] [--zoom Z] [--depth S]
 *            [--spheres N] [--lights L] [--seed K] [--device D] [--gpus G] [--list]
 *            [--out file.ppm|file.png] [--frames F] [--zoom-step dZ] [--scene file] [--save-scene file] [--accel]
 *
 * --gpus G     the frame is sharded by row strips over G GPUs of this box (include/rt_cuda_multi.h:
 *              one stream per device, NCCL max all-reduce + RGB8 all-gather); same bytes as 1 GPU.
 * --frames F   an animation with the scene resident on the device: frame f is rendered with
 *              zoom + f*dZ (--zoom-step).  The quantised frame f is copied back asynchronously
 *              (rt_cuda_readback_rgb8_async, pinned double buffer) and written to disk while the
 *              GPU renders frame f+1.  With F > 1 and a `%` in --out (e.g. frame_%03d.ppm) every
 *              frame is written, otherwise only the last.
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "rt_cuda.h"
#include "rt_cuda_multi.h"
#include "rt_scene.h"
#include "rt_png.h"

/* err_code.h:142-155 keeps print-and-exit as the HOST's policy */
static void check(int status, const char* what, rt_cuda_ctx* ctx = nullptr) {
  if (status != RT_CUDA_OK) {
    fprintf(stderr, "! %s: %s %s\n", what, rt_cuda_strerror(status), ctx ? rt_cuda_last_error(ctx) : "");
    exit(EXIT_FAILURE);
  }
}

static bool save_ppm(const unsigned char* rgb, const char* filename, unsigned width, unsigned height) {
  if (width == 0 || height == 0) { fprintf(stderr, "Can't save an empty image\n"); return false; }
  std::ofstream ofs(filename, std::ios::binary);
  if (ofs.fail()) { fprintf(stderr, "Can't open output file\n"); return false; }
  ofs << "P6\n" << width << " " << height << "\n255\n";
  ofs.write(reinterpret_cast<const char*>(rgb), (std::streamsize)width * height * 3);
  return !ofs.fail();
}

int main(int argc, char** argv) {
  unsigned width = 800, height = 600;        /* main.cpp:105-106 */
  float zoom = -4.f, alias = 3.f;            /* main.cpp:107-108 */
  int depth = 6;                             /* raytraceStack.h:10 */
  unsigned nSpheres = 0, nLights = 4, frames = 1;
  unsigned long long seed = 0;
  int device = 0, accel = 0, gpus = 0;
  float zoomStep = 0.f;                 /* --accel: optional two-level cluster filter (same frame, faster from ~768 spheres) */
  std::string out = "testPPM.ppm";           /* main.cpp:501 */
  std::string sceneFile, saveScene;

  for (int i = 1; i < argc; ++i) {
    auto need = [&](const char* flag) -> const char* {
      if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", flag); exit(EXIT_FAILURE); }
      return argv[++i];
    };
    if (!strcmp(argv[i], "--width")) width = (unsigned)atoi(need("--width"));
    else if (!strcmp(argv[i], "--height")) height = (unsigned)atoi(need("--height"));
    else if (!strcmp(argv[i], "--alias")) alias = (float)atof(need("--alias"));
    else if (!strcmp(argv[i], "--zoom")) zoom = (float)atof(need("--zoom"));
    else if (!strcmp(argv[i], "--depth")) depth = atoi(need("--depth"));
    else if (!strcmp(argv[i], "--spheres")) nSpheres = (unsigned)atoi(need("--spheres"));
    else if (!strcmp(argv[i], "--lights")) nLights = (unsigned)atoi(need("--lights"));
    else if (!strcmp(argv[i], "--seed")) seed = strtoull(need("--seed"), nullptr, 0);
    else if (!strcmp(argv[i], "--device")) device = atoi(need("--device"));
    else if (!strcmp(argv[i], "--frames")) frames = (unsigned)atoi(need("--frames"));
    else if (!strcmp(argv[i], "--zoom-step")) zoomStep = (float)atof(need("--zoom-step"));
    else if (!strcmp(argv[i], "--gpus")) gpus = atoi(need("--gpus"));
    else if (!strcmp(argv[i], "--out")) out = need("--out");
    else if (!strcmp(argv[i], "--scene")) sceneFile = need("--scene");
    else if (!strcmp(argv[i], "--save-scene")) saveScene = need("--save-scene");
    else if (!strcmp(argv[i], "--accel")) accel = 1;
    else if (!strcmp(argv[i], "--list")) {
      const int n = rt_cuda_device_count();
      printf("%d CUDA device(s)\n", n);
      for (int d = 0; d < n; ++d) {
        char buf[512];
        if (rt_cuda_device_info(d, buf, sizeof buf) == RT_CUDA_OK) printf("  %d: %s\n", d, buf);
      }
      return 0;
    } else {
      fprintf(stderr, "usage: %s [--width W] [--height H] [--alias A] [--zoom Z] [--depth S] "
                      "[--spheres N] [--lights L] [--seed K] [--device D] [--gpus G] [--frames F] [--zoom-step dZ] "
                      "[--out file.ppm|file.png] [--scene file] [--save-scene file] [--accel] [--list]\n",
              argv[0]);
      return EXIT_FAILURE;
    }
  }

  /* the scene: the reference literal, or synth(N, L, seed) */
  std::vector<rt_sphere> spheres;
  std::vector<rt_light> lights;
  if (!sceneFile.empty()) {
    rt_sphere* s = nullptr; rt_light* l = nullptr; unsigned ns = 0, nl = 0;
    if (rt_scene_load(sceneFile.c_str(), &s, &ns, &l, &nl) != 0) {
      fprintf(stderr, "cannot read scene file %s\n", sceneFile.c_str());
      return EXIT_FAILURE;
    }
    spheres.assign(s, s + ns); lights.assign(l, l + nl);
    rt_scene_free(s); rt_scene_free(l);
  } else if (nSpheres == 0) {
    spheres.resize(RT_SCENE_DEFAULT_SPHERES);
    lights.resize(RT_SCENE_DEFAULT_LIGHTS);
    rt_scene_default(spheres.data(), lights.data());
  } else {
    spheres.resize(nSpheres);
    lights.resize(nLights);
    if (rt_scene_synth(nSpheres, nLights, seed, spheres.data(), lights.data()) != 0) {
      fprintf(stderr, "bad synthetic scene parameters\n");
      return EXIT_FAILURE;
    }
  }

  if (!saveScene.empty() &&
      rt_scene_save(saveScene.c_str(), spheres.data(), (unsigned)spheres.size(), lights.data(), (unsigned)lights.size()) != 0) {
    fprintf(stderr, "cannot write scene file %s\n", saveScene.c_str());
    return EXIT_FAILURE;
  }

  const bool png = out.size() > 4 && out.compare(out.size() - 4, 4, ".png") == 0;      /* --out x.png: PNG instead of PPM */
  const bool perFrame = frames > 1 && out.find('%') != std::string::npos;
  auto write_frame = [&](const unsigned char* rgb, unsigned f) -> bool {
    char name[1024];
    if (perFrame) snprintf(name, sizeof name, out.c_str(), f); else snprintf(name, sizeof name, "%s", out.c_str());
    const bool ok = png ? rtpng::write_rgb8(name, rgb, width, height) : save_ppm(rgb, name, width, height);
    if (ok) printf("wrote %s\n", name);
    return ok;
  };
  const size_t frameBytes = (size_t)width * height * 3;
  if (frames == 0) frames = 1;

  if (gpus > 0) {
    /* ---- the frame sharded over `gpus` devices of this box ---- */
    rt_cuda_multi* m = nullptr;
    if (rt_cuda_multi_init(gpus, nullptr, &m) != RT_CUDA_OK) {
      fprintf(stderr, "! Opening %d CUDA devices failed (%d present)\n", gpus, rt_cuda_device_count());
      return EXIT_FAILURE;
    }
    auto mcheck = [&](int status, const char* what) {
      if (status != RT_CUDA_OK) {
        fprintf(stderr, "! %s: %s %s\n", what, rt_cuda_strerror(status), rt_cuda_multi_last_error(m));
        exit(EXIT_FAILURE);
      }
    };
    if (accel) mcheck(rt_cuda_multi_set_option(m, "accel", 1), "Selecting the accelerated mode");
    char info[512];
    for (int g = 0; g < gpus; ++g)
      if (rt_cuda_device_info(g, info, sizeof info) == RT_CUDA_OK) printf("%s\n", info);
    mcheck(rt_cuda_multi_upload_scene(m, spheres.data(), (unsigned)spheres.size(), lights.data(), (unsigned)lights.size()),
           "Copying the scene to the devices");
    std::vector<unsigned char> rgb(frameBytes);
    float maxColourValue = 1.f;
    for (unsigned f = 0; f < frames; ++f) {
      auto t0 = std::chrono::steady_clock::now();
      mcheck(rt_cuda_multi_render(m, width, height, zoom + (float)f * zoomStep, alias, depth, 0), "Enqueueing the frame");
      mcheck(rt_cuda_multi_synchronize(m), "Waiting for commands to finish");
      auto t1 = std::chrono::steady_clock::now();
      printf("Exec time: %.5f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
      mcheck(rt_cuda_multi_readback_rgb8(m, 0, rgb.data(), &maxColourValue), "Reading the image back");
      if ((perFrame || f + 1 == frames) && !write_frame(rgb.data(), f)) return EXIT_FAILURE;
    }
    unsigned long long rays = 0;
    float stepMs = 0.f;
    for (int g = 0; g < gpus; ++g) {
      rt_cuda_stats st;
      if (rt_cuda_get_stats(rt_cuda_multi_context(m, g), &st) == RT_CUDA_OK) {
        rays += st.rays;
        printf("  gpu %d: rows %u  trace kernel %.3f ms  rays %llu\n", g, st.local_rows, st.kernel_ms, (unsigned long long)st.rays);
      }
      float ms = 0.f;
      if (rt_cuda_multi_step_ms(m, g, &ms) == RT_CUDA_OK && ms > stepMs) stepMs = ms;
    }
    printf("%d GPUs: step %.3f ms (max over devices)  rays %llu (%.1f Mrays/s)  max colour %g\n", gpus, stepMs, rays,
           stepMs > 0 ? rays / stepMs / 1e3 : 0.0, maxColourValue);
    rt_cuda_multi_destroy(m);
    return 0;
  }

  rt_cuda_ctx* ctx = nullptr;
  check(rt_cuda_init(device, &ctx), "Opening the CUDA device");
  if (accel) check(rt_cuda_set_option(ctx, "accel", 1), "Selecting the accelerated mode");
  char info[512];
  if (rt_cuda_device_info(device, info, sizeof info) == RT_CUDA_OK) printf(" \n%s\n", info);
  check(rt_cuda_upload_scene(ctx, spheres.data(), (unsigned)spheres.size(), lights.data(),
                             (unsigned)lights.size()), "Copying the scene to the device", ctx);

  /* pinned double buffer: frame f is on its way to the host while frame f+1 renders */
  unsigned char* host[2] = {(unsigned char*)rt_cuda_host_alloc(frameBytes), (unsigned char*)rt_cuda_host_alloc(frameBytes)};
  if (!host[0] || !host[1]) { fprintf(stderr, "! cannot allocate pinned host memory\n"); return EXIT_FAILURE; }
  int ticket[2] = {-1, -1};
  bool failed = false;
  auto t_begin = std::chrono::steady_clock::now();
  for (unsigned f = 0; f <= frames && !failed; ++f) {
    if (f < frames) {
      /* same interval as main.cpp:353-369 for a single frame: launch -> finish */
      auto t0 = std::chrono::steady_clock::now();
      check(rt_cuda_render(ctx, width, height, zoom + (float)f * zoomStep, alias, depth), "Enqueueing kernel", ctx);
      if (frames == 1) {
        check(rt_cuda_synchronize(ctx), "Waiting for commands to finish", ctx);
        printf("Exec time: %.5f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
      /* quantise on the device with the frame's own maximum, copy 3 B/px back without waiting */
      check(rt_cuda_readback_rgb8_async(ctx, host[f & 1], 0.f, &ticket[f & 1]), "Reading the image back", ctx);
    }
    if (f > 0) {               /* frame f-1: wait for its copy, write it while the GPU works on frame f */
      const unsigned g = f - 1;
      check(rt_cuda_readback_wait(ctx, ticket[g & 1]), "Waiting for the image", ctx);
      if ((perFrame || g + 1 == frames) && !write_frame(host[g & 1], g)) failed = true;
    }
  }
  const double total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (frames > 1) printf("%u frames in %.3f ms (%.2f frames/s, render + readback + file output overlapped)\n", frames, total_ms,
                         frames * 1e3 / total_ms);
  rt_cuda_stats st;
  check(rt_cuda_get_stats(ctx, &st), "Reading statistics", ctx);
  const float maxColourValue = st.max_colour == 0.f ? 1.f : st.max_colour;
  printf("kernel %.3f ms  rays %llu (%.1f Mrays/s)  sphere tests %llu  max colour %g\n", st.kernel_ms,
         (unsigned long long)st.rays, st.kernel_ms > 0 ? st.rays / st.kernel_ms / 1e3 : 0.0,
         (unsigned long long)(st.rays * st.sph_num), maxColourValue);
  rt_cuda_host_free(host[0]); rt_cuda_host_free(host[1]);
  rt_cuda_destroy(ctx);
  return failed ? EXIT_FAILURE : 0;
}
