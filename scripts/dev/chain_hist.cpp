/* Development aid: distribution of the number of dependent queries per sample (the chain a lane works off)
 * on the bench frame, sub-sampled.  g++ -O2 -fopenmp -ffp-contract=off -mfma -Iinclude -Iraytracer-gamma_b200/csrc */
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "rt_core.cuh"
#include "rt_soa.h"
#include "rt_scene.h"
using namespace rtg;
int main(int argc, char** argv) {
  unsigned n = argc > 1 ? atoi(argv[1]) : 1024, W = argc > 2 ? atoi(argv[2]) : 7680, H = argc > 3 ? atoi(argv[3]) : 4320;
  float alias = argc > 4 ? atof(argv[4]) : 2.f; int S = argc > 5 ? atoi(argv[5]) : 8; unsigned step = argc > 6 ? atoi(argv[6]) : 16;
  std::vector<rt_sphere> sph(n); std::vector<rt_light> lg(4);
  rt_scene_synth(n, 4, 0, sph.data(), lg.data());
  std::vector<float4_> h; SceneLayout lay;
  build_scene_soa(sph.data(), n, lg.data(), 4, h, lay);
  const SceneView sc = scene_view(h.data(), lay);
  const Camera cam = make_camera(W, H, -4.f, alias, S, (int)sc.n);
  std::vector<unsigned> lens;
  std::vector<std::vector<unsigned>> perThread;
#pragma omp parallel
  {
    std::vector<unsigned> mine;
#pragma omp for schedule(dynamic, 1)
    for (long gy = 0; gy < (long)H; gy += step)
      for (unsigned gx = (gy / step * 7) % step; gx < W; gx += step)
        for (int si = 0; si < cam.nIter; ++si) for (int sj = 0; sj < cam.nIter; ++sj) {
          Slot s; Frame stack[RT_MAX_STACK]; Counters ctr; memset(&ctr, 0, sizeof ctr); memset(&s, 0, sizeof s);
          start_task(s, ctr, cam, gx, gy, 0, si, sj);
          unsigned q = 0;
          for (;;) {
            ++q;
            ShadowGeo sg;
            /* exact answers, no filter (same results) */
            if (s.kind == K_TRACE) { s.minT = 1000.f; s.hitIdx = -1; if (vdot(s.rayD, s.rayD) != 0.f) for (uint32_t i = 0; i < sc.n; ++i) resolve_trace(s.minT, s.hitIdx, s.qo, s.rayD, sc.geo[i], i); }
            else if (s.kind == K_SHADOW) { s.blocked = 0; shadow_geo(s, sc, sg); for (int k = 0; k < s.ndirs; ++k) for (uint32_t i = 0; i < sc.n; ++i) if (resolve_shadow(s.P, sg.d[k], sg.gap[k], sc.geo[i])) { s.blocked |= 1u << k; break; } }
            else { s.hitIdx = -1; for (uint32_t i = 0; i < sc.n; ++i) resolve_contain(s.hitIdx, s.qo, sc.geo[i], i); }
            if (advance(s, stack, ctr, sc, cam, &sg)) break;
          }
          mine.push_back(q);
        }
#pragma omp critical
    lens.insert(lens.end(), mine.begin(), mine.end());
  }
  std::sort(lens.begin(), lens.end());
  double sum = 0; for (unsigned v : lens) sum += v;
  printf("samples %zu mean %.2f median %u p90 %u p99 %u p99.9 %u p99.99 %u max %u\n", lens.size(), sum / lens.size(), lens[lens.size() / 2],
         lens[lens.size() * 9 / 10], lens[lens.size() * 99 / 100], lens[(size_t)(lens.size() * 0.999)], lens[(size_t)(lens.size() * 0.9999)], lens.back());
  /* share of all queries in samples longer than L */
  for (unsigned L : {16u, 32u, 64u, 128u, 256u, 512u}) { double a = 0; size_t c = 0; for (unsigned v : lens) if (v > L) { a += v; ++c; } printf("len > %3u: %.4f of samples, %.3f of queries\n", L, (double)c / lens.size(), a / sum); }
  return 0;
}
