/* hostsim.cpp — CPU lane simulator of the trace machine in rt_core.cuh.
 *
 * TEST INFRASTRUCTURE ONLY (built into tests/_build/libhostsim.so by
 * __graft_entry__.build()).  It runs the very same state machine, filter and
 * exact tests the CUDA kernel runs, one lane at a time, so the control flow of
 * rt_core.cuh can be checked against the oracle on a machine without a GPU.
 * It is not reachable from the product path (librt_cuda.so never links it).
 * Compile with -ffp-contract=off -mfma (fmaf must be a true fused operation). */
#include <stdint.h>
#include <string.h>
#include <vector>

#include "rt_core.cuh"
#include "rt_soa.h"

using namespace rtg;

extern "C" int hostsim_render(const rt_sphere* spheres, unsigned n, const rt_light* lights,
                              unsigned nl, unsigned W, unsigned H, float zoom, float alias, int S,
                              unsigned row_begin, unsigned row_count, unsigned row_step,
                              float* out, uint64_t* counters /* [8] or NULL */, int noFilter) {
  if (!out || W == 0 || H == 0 || S < 1 || S > RT_MAX_STACK) return -1;
  if (row_step == 0) row_step = 1;
  std::vector<float4_> h;
  SceneLayout lay;
  build_scene_soa(spheres, n, lights, nl, h, lay);
  const SceneView sc = scene_view(h.data(), lay);
  const Camera cam = make_camera(W, H, zoom, alias, S, (int)sc.n);
  uint64_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};

#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c[:8])
  for (long k = 0; k < (long)row_count; ++k) {
    const unsigned gy = row_begin + (unsigned)k * row_step;
    for (unsigned gx = 0; gx < W; ++gx) {
      Lane L;
      Frame stack[RT_MAX_STACK];
      memset(&L, 0, sizeof L);
      L.medium = (int)sc.n;
      float* px = out + ((size_t)k * W + gx) * 3;
      if (!start_pixel(L, cam, gx, gy, 0)) { px[0] = px[1] = px[2] = 0.f; continue; }
      for (;;) {
        c[6]++; c[7]++;
        if (noFilter) {
          if (L.qy.q != INFINITY)
            for (uint32_t i = 0; i < sc.n; ++i) resolve_candidate(L, sc, i);
        } else {
          for (uint32_t i = 0; i < sc.nPad; ++i)
            if (filter_pass(L.qy, sc.filt[i])) resolve_candidate(L, sc, i);
        }
        if (advance(L, sc, stack, cam)) {
          if (finish_sample(L, cam)) break;
        }
      }
      px[0] = L.acc.x; px[1] = L.acc.y; px[2] = L.acc.z;
      c[0] += L.ctr.rays; c[1] += L.ctr.shadow; c[2] += L.ctr.containQ; c[3] += L.ctr.containT;
      c[4] += L.ctr.exactTests; c[5] += L.ctr.samples;
    }
  }
  if (counters) memcpy(counters, c, sizeof c);
  return 0;
}
