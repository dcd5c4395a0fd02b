/* ref_opencl.cpp — runs the reference's OWN OpenCL kernel, unmodified, on whatever OpenCL GPU the box has.
 *
 * TEST / BASELINE INFRASTRUCTURE (never part of the product path).  Built by `make -C oracle opencl` into
 * oracle/_ref/ref_opencl only where /root/reference exists: the three source files the kernel consists of
 * (raytrace_kernel.cl and the vec.h / algebra.h it includes, /root/reference/raytracer_gamma/) are embedded
 * as data with `ld -r -b binary` straight from where they lie — nothing of them is copied into the repo —
 * and handed to the OpenCL compiler at run time exactly as main.cpp:236-263 does (JIT, no build options
 * beyond the include resolution).  The host side restates the launch of main.cpp:277-362: buffers, the 11
 * kernel arguments, a 1-D NDRange of W*H rounded up to the work-group size, clFinish, and the reference's
 * own timing interval ("Exec time": launch -> finish).  The kernel has no `gid < W*H` guard
 * (raytrace_kernel.cl:884,972), so the destination buffer is allocated for the padded range.
 *
 *   ref_opencl W H alias out.f32 [spheres.bin lights.bin]   -> prints "Exec time: X ms" (best of 3), writes W*H*3 floats
 *
 * Scene: the literal of main.cpp:113-168 (via librt_scene.so) unless raw Sphere / Light arrays are given.
 * Limits inherited from the kernel: sphNum and lgtNum <= work-group size, 48 B/sphere of __local memory,
 * RTSTACK_MAXSIZE 5, float Fresnel term (raytrace_kernel.cl:58,399-432).
 * Exit codes: 0 ok, 3 no OpenCL platform / GPU device (nothing measured), 1 any other failure. */
#define CL_TARGET_OPENCL_VERSION 120
#include <CL/cl.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rt_scene.h"

extern "C" {
extern const char _binary_raytrace_kernel_cl_start[], _binary_raytrace_kernel_cl_end[];
extern const char _binary_vec_h_start[], _binary_vec_h_end[];
extern const char _binary_algebra_h_start[], _binary_algebra_h_end[];
}

#define CK(err, what) do { if ((err) != CL_SUCCESS) { fprintf(stderr, "! %s: OpenCL error %d\n", what, (int)(err)); return 1; } } while (0)

static std::vector<char> slurp(const char* path) {
  std::vector<char> v;
  FILE* f = fopen(path, "rb");
  if (!f) return v;
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  v.resize((size_t)n);
  if (fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s W H alias out.f32 [spheres.bin lights.bin]\n", argv[0]); return 1; }
  const unsigned W = (unsigned)atoi(argv[1]), H = (unsigned)atoi(argv[2]);
  const float alias = (float)atof(argv[3]), zoom = -4.f;               /* main.cpp:107 */
  std::vector<rt_sphere> spheres(RT_SCENE_DEFAULT_SPHERES);
  std::vector<rt_light> lights(RT_SCENE_DEFAULT_LIGHTS);
  rt_scene_default(spheres.data(), lights.data());
  if (argc >= 7) {
    const std::vector<char> s = slurp(argv[5]), l = slurp(argv[6]);
    if (s.empty() || s.size() % sizeof(rt_sphere) || l.size() % sizeof(rt_light)) { fprintf(stderr, "! bad scene arrays\n"); return 1; }
    spheres.assign((const rt_sphere*)s.data(), (const rt_sphere*)(s.data() + s.size()));
    lights.assign((const rt_light*)l.data(), (const rt_light*)(l.data() + l.size()));
  }
  const cl_uint sphNum = (cl_uint)spheres.size(), lgtNum = (cl_uint)lights.size();

  cl_uint np = 0;
  if (clGetPlatformIDs(0, nullptr, &np) != CL_SUCCESS || np == 0) { printf("NO_OPENCL_PLATFORM\n"); return 3; }
  std::vector<cl_platform_id> plats(np);
  clGetPlatformIDs(np, plats.data(), nullptr);
  cl_device_id dev = nullptr;
  for (cl_uint i = 0; i < np && !dev; ++i)                              /* main.cpp:185-211: first GPU of the first platform that has one */
    if (clGetDeviceIDs(plats[i], CL_DEVICE_TYPE_GPU, 1, &dev, nullptr) != CL_SUCCESS) dev = nullptr;
  if (!dev) { printf("NO_OPENCL_GPU\n"); return 3; }
  char name[256] = {0}, ver[256] = {0};
  clGetDeviceInfo(dev, CL_DEVICE_NAME, sizeof name, name, nullptr);
  clGetDeviceInfo(dev, CL_DEVICE_VERSION, sizeof ver, ver, nullptr);
  size_t maxWg = 0; cl_ulong localMem = 0;
  clGetDeviceInfo(dev, CL_DEVICE_MAX_WORK_GROUP_SIZE, sizeof maxWg, &maxWg, nullptr);
  clGetDeviceInfo(dev, CL_DEVICE_LOCAL_MEM_SIZE, sizeof localMem, &localMem, nullptr);
  printf("OpenCL device: %s (%s), max work-group %zu, local memory %llu B\n", name, ver, maxWg, (unsigned long long)localMem);
  if (sphNum > maxWg || lgtNum > maxWg || sizeof(rt_sphere) * sphNum + sizeof(rt_light) * lgtNum > localMem) {
    printf("SCENE_TOO_LARGE_FOR_THE_REFERENCE_KERNEL (%u spheres: it stages the scene with one work-item per element into __local memory)\n", sphNum);
    return 3;
  }

  cl_int err;
  cl_context ctx = clCreateContext(nullptr, 1, &dev, nullptr, nullptr, &err); CK(err, "clCreateContext");
  cl_command_queue q = clCreateCommandQueue(ctx, dev, 0, &err); CK(err, "clCreateCommandQueue");

  /* the kernel source with its two headers, compiled and linked at run time (the reference passes -I <its folder>) */
  const char* ksrc = _binary_raytrace_kernel_cl_start; size_t klen = (size_t)(_binary_raytrace_kernel_cl_end - _binary_raytrace_kernel_cl_start);
  const char* vsrc = _binary_vec_h_start; size_t vlen = (size_t)(_binary_vec_h_end - _binary_vec_h_start);
  const char* asrc = _binary_algebra_h_start; size_t alen = (size_t)(_binary_algebra_h_end - _binary_algebra_h_start);
  cl_program pk = clCreateProgramWithSource(ctx, 1, &ksrc, &klen, &err); CK(err, "clCreateProgramWithSource");
  cl_program hdr[2] = {clCreateProgramWithSource(ctx, 1, &vsrc, &vlen, &err), clCreateProgramWithSource(ctx, 1, &asrc, &alen, &err)};
  const char* hdrNames[2] = {"vec.h", "algebra.h"};
  err = clCompileProgram(pk, 1, &dev, "", 2, hdr, hdrNames, nullptr, nullptr);
  if (err != CL_SUCCESS) {
    size_t n = 0; clGetProgramBuildInfo(pk, dev, CL_PROGRAM_BUILD_LOG, 0, nullptr, &n);
    std::string log(n, 0); clGetProgramBuildInfo(pk, dev, CL_PROGRAM_BUILD_LOG, n, &log[0], nullptr);
    fprintf(stderr, "! clCompileProgram failed (%d):\n%s\n", (int)err, log.c_str());
    return 1;
  }
  cl_program prog = clLinkProgram(ctx, 1, &dev, "", 1, &pk, nullptr, nullptr, &err); CK(err, "clLinkProgram");
  cl_kernel k = clCreateKernel(prog, "raytrace", &err); CK(err, "clCreateKernel");

  const size_t local = maxWg;                                            /* main.cpp:306-334 */
  const size_t count = (size_t)W * H;
  const size_t global = (count + local - 1) / local * local;
  cl_mem dS = clCreateBuffer(ctx, CL_MEM_READ_WRITE, sizeof(rt_sphere) * sphNum, nullptr, &err); CK(err, "clCreateBuffer spheres");
  cl_mem dL = clCreateBuffer(ctx, CL_MEM_READ_WRITE, sizeof(rt_light) * (lgtNum ? lgtNum : 1), nullptr, &err); CK(err, "clCreateBuffer lights");
  cl_mem dP = clCreateBuffer(ctx, CL_MEM_WRITE_ONLY, sizeof(rt_vec) * global, nullptr, &err); CK(err, "clCreateBuffer pixels");
  CK(clEnqueueWriteBuffer(q, dS, CL_TRUE, 0, sizeof(rt_sphere) * sphNum, spheres.data(), 0, nullptr, nullptr), "write spheres");
  if (lgtNum) CK(clEnqueueWriteBuffer(q, dL, CL_TRUE, 0, sizeof(rt_light) * lgtNum, lights.data(), 0, nullptr, nullptr), "write lights");
  cl_uint a = 0;                                                         /* main.cpp:339-349 */
  err  = clSetKernelArg(k, a++, sizeof(cl_mem), &dS);
  err |= clSetKernelArg(k, a++, sizeof(cl_uint), &sphNum);
  err |= clSetKernelArg(k, a++, sizeof(cl_mem), &dL);
  err |= clSetKernelArg(k, a++, sizeof(cl_uint), &lgtNum);
  err |= clSetKernelArg(k, a++, sizeof(cl_uint), &W);
  err |= clSetKernelArg(k, a++, sizeof(cl_uint), &H);
  err |= clSetKernelArg(k, a++, sizeof(cl_float), &zoom);
  err |= clSetKernelArg(k, a++, sizeof(cl_float), &alias);
  err |= clSetKernelArg(k, a++, sizeof(cl_mem), &dP);
  err |= clSetKernelArg(k, a++, sizeof(rt_sphere) * sphNum, nullptr);
  err |= clSetKernelArg(k, a++, sizeof(rt_light) * (lgtNum ? lgtNum : 1), nullptr);
  CK(err, "clSetKernelArg");

  double best = 1e30;
  for (int rep = 0; rep < 4; ++rep) {                                   /* first launch warms the JIT'd code up */
    auto t0 = std::chrono::steady_clock::now();                         /* main.cpp:353-369 */
    CK(clEnqueueNDRangeKernel(q, k, 1, nullptr, &global, &local, 0, nullptr, nullptr), "clEnqueueNDRangeKernel");
    CK(clFinish(q), "clFinish");
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rep > 0 && ms < best) best = ms;
  }
  printf("Exec time: %.5f ms\n", best);
  std::vector<float> px(count * 3);
  CK(clEnqueueReadBuffer(q, dP, CL_TRUE, 0, sizeof(rt_vec) * count, px.data(), 0, nullptr, nullptr), "read pixels");   /* main.cpp:460 */
  FILE* f = fopen(argv[4], "wb");
  if (!f || fwrite(px.data(), sizeof(float), px.size(), f) != px.size()) { fprintf(stderr, "! cannot write %s\n", argv[4]); return 1; }
  fclose(f);
  clReleaseMemObject(dS); clReleaseMemObject(dL); clReleaseMemObject(dP);
  clReleaseKernel(k); clReleaseProgram(prog); clReleaseProgram(pk); clReleaseProgram(hdr[0]); clReleaseProgram(hdr[1]);
  clReleaseCommandQueue(q); clReleaseContext(ctx);
  return 0;
}
