/* main.cpp — the host program of raytracer-gamma with the OpenCL plumbing
 * (reference main.cpp:182-350, 456-489, device_info.cpp, device_picker.h) replaced by
 * the C-ABI of include/rt_cuda.h.  Call sequence and output are the reference's:
 * build the scene (main.cpp:105-168), render, print "Exec time", read the float
 * pixels back, normalise by the global maximum (algebra.h:68-91) and write a
 * binary PPM (main.cpp:43-91).  Unlike the reference it takes its constants from
 * the command line and does not wait for a key press (main.cpp:505).
 *
 *   rt_gamma [--width W] [--height H] [--alias A] [--zoom Z] [--depth S]
 *            [--spheres N] [--lights L] [--seed K] [--device D] [--list]
 *            [--out file.ppm|file.png] [--frames F] [--scene file] [--save-scene file] [--accel]
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "rt_cuda.h"
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
  int device = 0, accel = 0;                 /* --accel: optional two-level cluster filter (same frame, faster from ~768 spheres) */
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
                      "[--spheres N] [--lights L] [--seed K] [--device D] [--frames F] [--out file.ppm] [--scene file] "
                      "[--save-scene file] [--accel] [--list]\n",
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

  rt_cuda_ctx* ctx = nullptr;
  check(rt_cuda_init(device, &ctx), "Opening the CUDA device");
  if (accel) check(rt_cuda_set_option(ctx, "accel", 1), "Selecting the accelerated mode");
  char info[512];
  if (rt_cuda_device_info(device, info, sizeof info) == RT_CUDA_OK) printf(" \n%s\n", info);
  check(rt_cuda_upload_scene(ctx, spheres.data(), (unsigned)spheres.size(), lights.data(),
                             (unsigned)lights.size()), "Copying the scene to the device", ctx);

  std::vector<unsigned char> rgb((size_t)width * height * 3);
  float maxColourValue = 1.f;
  for (unsigned f = 0; f < frames; ++f) {
    /* same interval as main.cpp:353-369: launch -> finish */
    auto t0 = std::chrono::steady_clock::now();
    check(rt_cuda_render(ctx, width, height, zoom, alias, depth), "Enqueueing kernel", ctx);
    check(rt_cuda_synchronize(ctx), "Waiting for commands to finish", ctx);
    auto t1 = std::chrono::steady_clock::now();
    printf("Exec time: %.5f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
    /* quantise on the device with the frame's own maximum, copy 3 B/px back */
    check(rt_cuda_readback_rgb8(ctx, rgb.data(), 0.f), "Reading the image back", ctx);
  }
  rt_cuda_stats st;
  check(rt_cuda_get_stats(ctx, &st), "Reading statistics", ctx);
  maxColourValue = st.max_colour == 0.f ? 1.f : st.max_colour;
  printf("kernel %.3f ms  rays %llu (%.1f Mrays/s)  sphere tests %llu  max colour %g\n", st.kernel_ms,
         (unsigned long long)st.rays, st.kernel_ms > 0 ? st.rays / st.kernel_ms / 1e3 : 0.0,
         (unsigned long long)(st.rays * st.sph_num), maxColourValue);
  rt_cuda_destroy(ctx);

  const bool png = out.size() > 4 && out.compare(out.size() - 4, 4, ".png") == 0;      /* --out x.png: PNG instead of PPM */
  if (!(png ? rtpng::write_rgb8(out.c_str(), rgb.data(), width, height) : save_ppm(rgb.data(), out.c_str(), width, height)))
    return EXIT_FAILURE;
  printf("wrote %s\n", out.c_str());
  return 0;
}
