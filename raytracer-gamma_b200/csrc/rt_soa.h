/* rt_soa.h — host-side conversion of the reference's AoS scene (sphere.h:9-14,
 * raytracer.h:20-25) into the SoA records the kernels read (rt_core.cuh SceneView).
 * Shared by the shim (rt_shim.cu) and the CPU lane simulator used in tests. */
#ifndef RT_SOA_H
#define RT_SOA_H

#include <math.h>
#include <cmath>
#include <vector>
#include "rt_types.h"
#include "rt_core.cuh"

namespace rtg {

/* largest float <= v */
static inline float float_round_down(double v) {
  float f = (float)v;
  if ((double)f > v) f = nextafterf(f, -INFINITY);
  return f;
}

struct SceneLayout {
  uint32_t n, nPad, nl;
  size_t offFilt, offGeo, offMatA, offMatB, offLpos, offLcol, total;   /* in records */
};

static inline SceneLayout scene_layout(uint32_t n, uint32_t nl) {
  SceneLayout L;
  L.n = n; L.nl = nl; L.nPad = (n + 31u) & ~31u;
  const size_t nGeo = n ? n : 1, nLgt = nl ? nl : 1;
  L.offFilt = 0;
  L.offGeo = L.offFilt + L.nPad;
  L.offMatA = L.offGeo + nGeo;
  L.offMatB = L.offMatA + (n + 1);
  L.offLpos = L.offMatB + (n + 1);
  L.offLcol = L.offLpos + nLgt;
  L.total = L.offLcol + nLgt;
  return L;
}

static inline SceneView scene_view(const float4_* base, const SceneLayout& L) {
  SceneView v;
  v.filt = base + L.offFilt; v.geo = base + L.offGeo;
  v.matA = base + L.offMatA; v.matB = base + L.offMatB;
  v.lpos = base + L.offLpos; v.lcol = base + L.offLcol;
  v.n = L.n; v.nPad = L.nPad; v.nl = L.nl;
  return v;
}

static inline void build_scene_soa(const rt_sphere* spheres, uint32_t n, const rt_light* lights,
                                   uint32_t nl, std::vector<float4_>& h, SceneLayout& L) {
  L = scene_layout(n, nl);
  h.assign(L.total, float4_{0.f, 0.f, 0.f, 0.f});
  float4_* filt = h.data() + L.offFilt;
  float4_* geo = h.data() + L.offGeo;
  float4_* matA = h.data() + L.offMatA;
  float4_* matB = h.data() + L.offMatB;
  float4_* lpos = h.data() + L.offLpos;
  float4_* lcol = h.data() + L.offLcol;
  const double kappa = (double)RT_KAPPA;
  for (uint32_t i = 0; i < L.nPad; ++i) {
    if (i < n) {
      const rt_sphere& s = spheres[i];
      const double cx = s.pos.x, cy = s.pos.y, cz = s.pos.z, r = s.radius;
      const double cc = cx * cx + cy * cy + cz * cz, rr = r * r;
      /* see "Filter" in rt_core.cuh: certain-miss threshold with all rounding slack folded in */
      const double w = (cc - rr) - kappa * (cc + rr) - 2.5e-6 * fabs(r) - 1e-11;
      const float wf = float_round_down(w);
      const bool finite = std::isfinite(s.pos.x) && std::isfinite(s.pos.y) && std::isfinite(s.pos.z) &&
                          std::isfinite(s.radius) && std::isfinite(wf);
      /* geometry the filter cannot represent is ALWAYS a candidate (w = -inf, finite products):
       * it then goes through the reference's exact expressions for every query */
      filt[i] = finite ? float4_{s.pos.x, s.pos.y, s.pos.z, wf} : float4_{0.f, 0.f, 0.f, -INFINITY};
      geo[i] = float4_{s.pos.x, s.pos.y, s.pos.z, s.radius};
      matA[i] = float4_{s.material.matteColour.x, s.material.matteColour.y,
                        s.material.matteColour.z, s.material.opacity};
      matB[i] = float4_{s.material.glossColour.x, s.material.glossColour.y,
                        s.material.glossColour.z, s.material.refractiveIndex};
    } else {
      filt[i] = float4_{0.f, 0.f, 0.f, INFINITY};   /* padding: never flags */
    }
  }
  /* the ambient medium (main.cpp:423-426, raytracer.h:694-697): black, index 1, and the
   * uninitialised opacity pinned to 0 (SURVEY.md §7 hard part 3) */
  matA[n] = float4_{0.f, 0.f, 0.f, 0.f};
  matB[n] = float4_{0.f, 0.f, 0.f, 1.00f};
  for (uint32_t l = 0; l < nl; ++l) {
    lpos[l] = float4_{lights[l].pos.x, lights[l].pos.y, lights[l].pos.z, 0.f};
    lcol[l] = float4_{lights[l].col.x, lights[l].col.y, lights[l].col.z, 0.f};
  }
}

}  // namespace rtg
#endif
