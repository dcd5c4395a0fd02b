/* fuzz_filter.cu — GPU fuzz of the per-sphere FMA filter of rt_core.cuh against the exact tests.
 *
 * TEST INFRASTRUCTURE ONLY (built into tests/_build/libfuzz.so by __graft_entry__.build_fuzz();
 * never linked into the product).  The filter may only ever say "certain miss" for a (query, sphere)
 * pair the reference's exact expressions (raytracer.h:81-141, :259-266) also reject; the bound behind
 * it (rt_core.cuh "Filter") is a first-order argument, so this hunts for violations where it is
 * tightest: coordinates up to 10^4, radii over six decades, rays aimed at the silhouette of the sphere
 * with a relative miss distance from 10^-8 to 10^-1 on either side, origins far from, near and inside
 * the sphere, unit (shadow-style, make_dir_unit), unnormalised (refracted-style, make_dir_fast) and
 * probe-point (containment) queries.  Every thread owns one sphere and fires `rays` queries at it. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "rt_types.h"
#include "rt_core.cuh"
#include "rt_soa.h"

using namespace rtg;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct Rng {
  uint64_t s;
  __device__ float u() { s = mix64(s); return (float)((s >> 40) & 0xFFFFFFu) * (1.f / 16777216.f); }          /* [0,1) */
  __device__ float sym() { return 2.f * u() - 1.f; }
  __device__ float logu(float lo, float hi) { return exp10f(lo + (hi - lo) * u()); }                       /* 10^U(lo,hi) */
  __device__ V3 dir() {
    for (;;) {
      const V3 v = mk(sym(), sym(), sym());
      const float n = v.x * v.x + v.y * v.y + v.z * v.z;
      if (n > 1e-4f && n <= 1.f) { const float s = rsqrtf(n); return mk(v.x * s, v.y * s, v.z * s); }
    }
  }
};

/* counters: [0] ray tests, [1] ray candidates (filter could not rule out), [2] exact ray hits,
 * [3] RAY VIOLATIONS (exact hit of a filtered-out pair), [4] rays the filter refused (not filterable),
 * [5] probe tests, [6] probe candidates, [7] probes inside, [8] PROBE VIOLATIONS */
__global__ void fuzz_kernel(const float4_* __restrict__ filt, const float4_* __restrict__ geo, uint32_t nSpheres,
                            uint32_t rays, uint64_t seed, float maxCoordLog, unsigned long long* out) {
  unsigned long long c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (uint32_t si = blockIdx.x * blockDim.x + threadIdx.x; si < nSpheres; si += gridDim.x * blockDim.x) {
    const float4_ f = filt[si], g = geo[si];
    const V3 ctr = mk(g.x, g.y, g.z);
    const float r = g.w;
    Rng rng{mix64(seed ^ ((uint64_t)si << 20))};
    for (uint32_t j = 0; j < rays; ++j) {
      /* origin: at a distance of 10^U(-3, 4.3) radii-ish units from the centre, or anywhere in the box */
      V3 o;
      const float pick = rng.u();
      if (pick < 0.5f) { const V3 w = rng.dir(); const float dist = r * rng.logu(-2.f, 3.f); o = mk(ctr.x + w.x * dist, ctr.y + w.y * dist, ctr.z + w.z * dist); }
      else if (pick < 0.9f) { const float m = rng.logu(-1.f, maxCoordLog); const V3 w = rng.dir(); o = mk(w.x * m, w.y * m, w.z * m); }
      else o = mk(0.f, 0.f, 0.f);
      /* aim at the silhouette: a point at distance r (1 + eps) from the centre, perpendicular to the view line */
      const V3 view = mk(ctr.x - o.x, ctr.y - o.y, ctr.z - o.z);
      V3 u = rng.dir();
      const float vv = view.x * view.x + view.y * view.y + view.z * view.z;
      if (vv > 0.f) {
        const float k = (u.x * view.x + u.y * view.y + u.z * view.z) / vv;
        u = mk(u.x - k * view.x, u.y - k * view.y, u.z - k * view.z);
        const float un = u.x * u.x + u.y * u.y + u.z * u.z;
        if (un > 0.f) { const float s = rsqrtf(un); u = mk(u.x * s, u.y * s, u.z * s); }
      }
      const float eps = (rng.u() < 0.5f ? -1.f : 1.f) * rng.logu(-8.f, -1.f);
      const float rr = r * (1.f + eps);
      const V3 target = mk(ctr.x + u.x * rr, ctr.y + u.y * rr, ctr.z + u.z * rr);
      V3 d = vsub(target, o);
      const bool unitStyle = rng.u() < 0.5f;
      DirQ D;
      bool live;
      if (unitStyle) { d = vunit(d); make_dir_unit(D, o, d); live = true; }                  /* shadow rays, primary / reflected rays */
      else { const float sc = rng.logu(-1.f, 1.f); d = vscale(sc, vunit(d)); live = make_dir_fast(D, o, d); }   /* refracted rays */
      const OriginQ O = make_origin(o);
      c[0]++;
      if (live && origin_filterable(O) && dir_filterable(D)) {
        const float v = filter_ray(O, D, filter_ch(O, f), f);
        const bool miss = v < 0.f;
        const float t = ray_sphere_t(g, o, d);
        if (!miss) c[1]++;
        if (t > 0.f) { c[2]++; if (miss) c[3]++; }
      } else {
        c[4]++;
      }
      /* containment probe at distance r (1 + eps') from the centre */
      const float eps2 = (rng.u() < 0.5f ? -1.f : 1.f) * rng.logu(-8.f, -1.f);
      const V3 w2 = rng.dir();
      const float r2 = r * (1.f + eps2);
      const V3 pnt = mk(ctr.x + w2.x * r2, ctr.y + w2.y * r2, ctr.z + w2.z * r2);
      const OriginQ P = make_origin(pnt);
      c[5]++;
      if (origin_filterable(P)) {
        const bool out_ = filter_point(P, filter_ch(P, f)) < 0.f;
        const bool in = contains_exact(g, pnt);
        if (!out_) c[6]++;
        if (in) { c[7]++; if (out_) c[8]++; }
      }
    }
  }
  for (int i = 0; i < 9; ++i) {
    unsigned long long v = c[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[i], v);
  }
}

extern "C" int fuzz_filter(uint64_t seed, uint32_t nSpheres, uint32_t rays, float maxCoordLog, float rLogLo, float rLogHi,
                           uint64_t out[9]) {
  /* spheres on the host (double-precision record arithmetic of rt_soa.h, exactly what the shim uploads) */
  std::vector<float4_> filt(nSpheres), geo(nSpheres);
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345;
  auto u = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)((s >> 40) & 0xFFFFFFu) / 16777216.0; };
  for (uint32_t i = 0; i < nSpheres; ++i) {
    const double m = (u() < 0.1) ? 0.0 : pow(10.0, -1.0 + (maxCoordLog + 1.0) * u());
    double x = 2 * u() - 1, y = 2 * u() - 1, z = 2 * u() - 1;
    const double n = sqrt(x * x + y * y + z * z) + 1e-30;
    const float cx = (float)(m * x / n), cy = (float)(m * y / n), cz = (float)(m * z / n);
    const float r = (float)pow(10.0, rLogLo + (rLogHi - rLogLo) * u());
    filt[i] = filter_record(cx, cy, cz, r);
    geo[i] = float4_{cx, cy, cz, r};
  }
  float4_ *dF = nullptr, *dG = nullptr;
  unsigned long long* dO = nullptr;
  if (cudaMalloc(&dF, nSpheres * sizeof(float4_)) != cudaSuccess || cudaMalloc(&dG, nSpheres * sizeof(float4_)) != cudaSuccess ||
      cudaMalloc(&dO, 9 * sizeof(unsigned long long)) != cudaSuccess)
    return -1;
  cudaMemcpy(dF, filt.data(), nSpheres * sizeof(float4_), cudaMemcpyHostToDevice);
  cudaMemcpy(dG, geo.data(), nSpheres * sizeof(float4_), cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, 9 * sizeof(unsigned long long));
  fuzz_kernel<<<148 * 8, 256>>>(dF, dG, nSpheres, rays, seed, maxCoordLog, dO);
  const cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[9];
  cudaMemcpy(h, dO, sizeof h, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 9; ++i) out[i] = h[i];
  cudaFree(dF); cudaFree(dG); cudaFree(dO);
  if (e != cudaSuccess) { fprintf(stderr, "fuzz_filter: %s\n", cudaGetErrorString(e)); return -2; }
  return 0;
}

/* sin_from_cos (rt_core.cuh) against the IEEE double square root it replaces, for EVERY float c in [0, 1)
 * (c and -c square to the same value).  out[0] = values checked, out[1] = MISMATCHES (must be 0). */
__global__ void sqrt_check_kernel(unsigned long long* out) {
  unsigned long long n = 0, bad = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t bits = blockIdx.x * blockDim.x + threadIdx.x; bits < 0x3F800000u; bits += stride) {
    const float c = __uint_as_float(bits);
    const float want = (float)__dsqrt_rn(__dadd_rn(1.0, -(double)__fmul_rn(c, c)));
    const float got = sin_from_cos(c);
    ++n;
    if (__float_as_uint(want) != __float_as_uint(got)) ++bad;
  }
  for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xFFFFFFFFu, n, o); bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], n); if (bad) atomicAdd(&out[1], bad); }
}
extern "C" int fuzz_sin_from_cos(uint64_t out[2]) {
  unsigned long long* dO = nullptr;
  if (cudaMalloc(&dO, 2 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  cudaMemset(dO, 0, 2 * sizeof(unsigned long long));
  sqrt_check_kernel<<<148 * 16, 256>>>(dO);
  const cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, dO, sizeof h, cudaMemcpyDeviceToHost);
  out[0] = h[0]; out[1] = h[1];
  cudaFree(dO);
  if (e != cudaSuccess) { fprintf(stderr, "fuzz_sin_from_cos: %s\n", cudaGetErrorString(e)); return -2; }
  return 0;
}
