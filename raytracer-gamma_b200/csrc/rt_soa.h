/* rt_soa.h — host-side conversion of the reference's AoS scene (sphere.h:9-14,
 * raytracer.h:20-25) into the SoA records the kernels read (rt_core.cuh SceneView).
 * Shared by the shim (rt_shim.cu) and the CPU lane simulator used in tests. */
#ifndef RT_SOA_H
#define RT_SOA_H

#include <math.h>
#include <algorithm>
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
  uint32_t n, nPad, nl, nc, ncPad;
  size_t offFilt, offGeo, offMatA, offMatB, offLpos, offLcol, offCfilt, offMfilt, offMidx, total;   /* in records */
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
  /* cluster form: ceil(n / RT_CLUSTER) clusters plus one per started group of non-finite spheres
   * is at most ceil(n / RT_CLUSTER) + 1; the builder fixes nc, the layout reserves the bound */
  const uint32_t ncMax = n ? (n + RT_CLUSTER - 1) / RT_CLUSTER + 1 : 0;
  L.nc = 0; L.ncPad = (ncMax + 15u) & ~15u;
  L.offCfilt = L.offLcol + nLgt;
  L.offMfilt = L.offCfilt + L.ncPad;
  L.offMidx = L.offMfilt + (size_t)L.ncPad * RT_CLUSTER;
  L.total = L.offMidx + ((size_t)L.ncPad * RT_CLUSTER * sizeof(unsigned short) + 15) / 16;
  return L;
}

static inline SceneView scene_view(const float4_* base, const SceneLayout& L) {
  SceneView v;
  v.filt = base + L.offFilt; v.geo = base + L.offGeo;
  v.matA = base + L.offMatA; v.matB = base + L.offMatB;
  v.lpos = base + L.offLpos; v.lcol = base + L.offLcol;
  v.n = L.n; v.nPad = L.nPad; v.nl = L.nl;
  v.cfilt = base + L.offCfilt; v.mfilt = base + L.offMfilt;
  v.midx = reinterpret_cast<const unsigned short*>(base + L.offMidx);
  v.nc = L.nc; v.ncPad = L.ncPad;
  return v;
}

/* Filter record of sphere (c, r) — see "Filter" in rt_core.cuh: certain-miss threshold with
 * all rounding slack folded in.  Geometry the filter cannot represent is ALWAYS a candidate
 * (w = -inf, finite products): it then goes through the reference's exact expressions. */
static inline float4_ filter_record(float px, float py, float pz, float radius) {
  const double kappa = (double)RT_KAPPA;
  const double cx = px, cy = py, cz = pz, r = radius;
  const double cc = cx * cx + cy * cy + cz * cz, rr = r * r;
  const double w = (cc - rr) - kappa * (cc + rr) - 2.5e-6 * fabs(r) - 1e-11;
  const float wf = float_round_down(w);
  const bool finite = std::isfinite(px) && std::isfinite(py) && std::isfinite(pz) &&
                      std::isfinite(radius) && std::isfinite(wf);
  return finite ? float4_{px, py, pz, wf} : float4_{0.f, 0.f, 0.f, -INFINITY};
}
static inline bool record_always(const float4_& f) { return f.w == -INFINITY; }

/* Spatial clusters of RT_CLUSTER spheres and their bounding records ("Cluster filter" in
 * rt_core.cuh).  Median splits along the widest axis until a node holds one cluster; spheres
 * the filter cannot represent go to clusters that are always candidates. */
static inline void build_clusters(const rt_sphere* spheres, uint32_t n, const float4_* filt,
                                  float4_* cfilt, float4_* mfilt, unsigned short* midx,
                                  uint32_t& nc, uint32_t ncPad) {
  std::vector<uint32_t> fin, odd;
  for (uint32_t i = 0; i < n; ++i) (record_always(filt[i]) ? odd : fin).push_back(i);
  std::vector<std::pair<uint32_t, uint32_t>> leaves;          /* [lo, hi) ranges of `fin` */
  std::vector<std::pair<uint32_t, uint32_t>> todo;
  if (!fin.empty()) todo.push_back({0u, (uint32_t)fin.size()});
  while (!todo.empty()) {
    const uint32_t lo = todo.back().first, hi = todo.back().second;
    todo.pop_back();
    const uint32_t m = hi - lo;
    if (m <= RT_CLUSTER) { leaves.push_back({lo, hi}); continue; }
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t k = lo; k < hi; ++k) {
      const float c[3] = {spheres[fin[k]].pos.x, spheres[fin[k]].pos.y, spheres[fin[k]].pos.z};
      for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], c[a]); mx[a] = std::max(mx[a], c[a]); }
    }
    int ax = 0;
    if (mx[1] - mn[1] > mx[ax] - mn[ax]) ax = 1;
    if (mx[2] - mn[2] > mx[ax] - mn[ax]) ax = 2;
    const uint32_t half = ((m / 2 + RT_CLUSTER - 1) / RT_CLUSTER) * RT_CLUSTER;   /* < m because m > RT_CLUSTER */
    auto key = [&](uint32_t i) { const rt_vec& p = spheres[i].pos; return ax == 0 ? p.x : ax == 1 ? p.y : p.z; };
    std::nth_element(fin.begin() + lo, fin.begin() + lo + half, fin.begin() + hi,
                     [&](uint32_t a, uint32_t b) { const float ka = key(a), kb = key(b); return ka < kb || (ka == kb && a < b); });
    todo.push_back({lo + half, hi});
    todo.push_back({lo, lo + half});
  }
  nc = 0;
  auto emit = [&](const uint32_t* ids, uint32_t m, bool always) {
    float4_ rec = float4_{0.f, 0.f, 0.f, -INFINITY};
    if (!always) {
      /* centre: box of the member spheres, then a few steps towards the member that sets the radius */
      double lo3[3] = {INFINITY, INFINITY, INFINITY}, hi3[3] = {-INFINITY, -INFINITY, -INFINITY};
      for (uint32_t k = 0; k < m; ++k) {
        const rt_sphere& s = spheres[ids[k]];
        const double c[3] = {s.pos.x, s.pos.y, s.pos.z}, ar = fabs((double)s.radius);
        for (int a = 0; a < 3; ++a) { lo3[a] = std::min(lo3[a], c[a] - ar); hi3[a] = std::max(hi3[a], c[a] + ar); }
      }
      double C[3] = {0.5 * (lo3[0] + hi3[0]), 0.5 * (lo3[1] + hi3[1]), 0.5 * (lo3[2] + hi3[2])};
      auto reach = [&](const double* ctr, uint32_t* far) {
        double best = -1.0;
        for (uint32_t k = 0; k < m; ++k) {
          const rt_sphere& s = spheres[ids[k]];
          const double dx = s.pos.x - ctr[0], dy = s.pos.y - ctr[1], dz = s.pos.z - ctr[2];
          const double v = sqrt(dx * dx + dy * dy + dz * dz) + fabs((double)s.radius);
          if (v > best) { best = v; if (far) *far = k; }
        }
        return best;
      };
      for (int it = 0; it < 32; ++it) {
        uint32_t far = 0;
        const double R0 = reach(C, &far);
        const rt_sphere& s = spheres[ids[far]];
        const double T[3] = {C[0] + 0.05 * (s.pos.x - C[0]), C[1] + 0.05 * (s.pos.y - C[1]), C[2] + 0.05 * (s.pos.z - C[2])};
        if (reach(T, nullptr) < R0) { C[0] = T[0]; C[1] = T[1]; C[2] = T[2]; } else break;
      }
      const float Cf[3] = {(float)C[0], (float)C[1], (float)C[2]};      /* the stored centre; bounds are taken to IT */
      double R2 = 0.0;
      for (uint32_t k = 0; k < m; ++k) {
        const rt_sphere& s = spheres[ids[k]];
        const double dx = (double)s.pos.x - Cf[0], dy = (double)s.pos.y - Cf[1], dz = (double)s.pos.z - Cf[2];
        const double D2 = dx * dx + dy * dy + dz * dz, D = sqrt(D2) * (1.0 + 1e-15), ar = fabs((double)s.radius);
        const double cc = (double)s.pos.x * s.pos.x + (double)s.pos.y * s.pos.y + (double)s.pos.z * s.pos.z;
        const double v = (ar + D) * (ar + D) + D2 / 32.0 + (double)RT_KAPPA2 * (cc + ar * ar) + 1e-4 * ar + 1e-10;
        R2 = std::max(R2, v);
      }
      float Rf = (float)sqrt(R2 * (1.0 + 1e-12));
      Rf = nextafterf(nextafterf(Rf, INFINITY), INFINITY);
      rec = filter_record(Cf[0], Cf[1], Cf[2], Rf);
    }
    cfilt[nc] = rec;
    for (uint32_t k = 0; k < RT_CLUSTER; ++k) {
      mfilt[(size_t)nc * RT_CLUSTER + k] = (k < m) ? filt[ids[k]] : float4_{0.f, 0.f, 0.f, INFINITY};
      midx[(size_t)nc * RT_CLUSTER + k] = (k < m) ? (unsigned short)ids[k] : (unsigned short)0x3FFFu;
    }
    ++nc;
  };
  for (const auto& lf : leaves) {
    std::sort(fin.begin() + lf.first, fin.begin() + lf.second);
    emit(fin.data() + lf.first, lf.second - lf.first, false);
  }
  for (uint32_t k = 0; k < odd.size(); k += RT_CLUSTER)
    emit(odd.data() + k, std::min<uint32_t>(RT_CLUSTER, (uint32_t)odd.size() - k), true);
  for (uint32_t c = nc; c < ncPad; ++c) {
    cfilt[c] = float4_{0.f, 0.f, 0.f, INFINITY};          /* padding: never flags */
    for (uint32_t k = 0; k < RT_CLUSTER; ++k) {
      mfilt[(size_t)c * RT_CLUSTER + k] = float4_{0.f, 0.f, 0.f, INFINITY};
      midx[(size_t)c * RT_CLUSTER + k] = (unsigned short)0x3FFFu;
    }
  }
}

static inline void build_scene_soa(const rt_sphere* spheres, uint32_t n, const rt_light* lights,
                                   uint32_t nl, std::vector<float4_>& h, SceneLayout& L,
                                   bool withClusters = true) {
  L = scene_layout(n, nl);
  h.assign(L.total, float4_{0.f, 0.f, 0.f, 0.f});
  float4_* filt = h.data() + L.offFilt;
  float4_* geo = h.data() + L.offGeo;
  float4_* matA = h.data() + L.offMatA;
  float4_* matB = h.data() + L.offMatB;
  float4_* lpos = h.data() + L.offLpos;
  float4_* lcol = h.data() + L.offLcol;
  for (uint32_t i = 0; i < L.nPad; ++i) {
    if (i < n) {
      const rt_sphere& s = spheres[i];
      filt[i] = filter_record(s.pos.x, s.pos.y, s.pos.z, s.radius);
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
  if (withClusters)
    build_clusters(spheres, n, filt, h.data() + L.offCfilt, h.data() + L.offMfilt,
                   reinterpret_cast<unsigned short*>(h.data() + L.offMidx), L.nc, L.ncPad);
}

}  // namespace rtg
#endif
