/* hostsim.cpp — CPU simulator of the trace machine in rt_core.cuh.
 *
 * TEST INFRASTRUCTURE ONLY (built into tests/_build/libhostsim.so by
 * __graft_entry__.build()).  It runs the very same slot state machine, filter and
 * exact tests the CUDA kernel runs, one pixel at a time, so the control flow of
 * rt_core.cuh can be checked against the oracle on a machine without a GPU.
 * It is not reachable from the product path (librt_cuda.so never links it).
 * Compile with -ffp-contract=off -mfma (fmaf must be a true fused operation). */
#include <stdint.h>
#include <string.h>
#include <vector>

#include "rt_core.cuh"
#include "rt_soa.h"

using namespace rtg;

/* The accelerated mode's answer (rt_kernels.cuh pass_*_accel): cluster records first, then the
 * members of the clusters that were not ruled out, then the exact tests with explicit index
 * tie-breaks.  With `audit` every sphere is ALSO tested exactly, and a sphere the exact test
 * accepts inside a cluster the cluster filter ruled out counts as a violation (there must be none). */
static void answer_accel(Slot& s, const SceneView& sc, Counters& ctr, ShadowGeo& sg, bool audit,
                         const std::vector<uint32_t>& clusterOf, uint64_t& violations, uint64_t& clusterTests) {
  const V3 org = (s.kind == K_SHADOW) ? s.P : s.qo;
  const OriginQ O = make_origin(org);
  const OriginQ OC = cluster_origin(O);
  const bool ofil = origin_filterable(O);
  std::vector<unsigned char> flagged(sc.nc);
  auto brute = [&](auto&& fn) { for (uint32_t i = 0; i < sc.n; ++i) fn(i); };
  auto culled = [&](const DirQ* D, auto&& fn) {
    for (uint32_t c = 0; c < sc.nc; ++c) {
      const float ch = filter_ch(OC, sc.cfilt[c]);
      const float v = D ? filter_ray(OC, *D, ch, sc.cfilt[c]) : filter_point(OC, ch);
      flagged[c] = !(v < 0.f);
      ++clusterTests;
      if (!flagged[c]) continue;
      for (uint32_t m = 0; m < RT_CLUSTER; ++m) {
        const float4_ rec = sc.mfilt[c * RT_CLUSTER + m];
        const float chm = filter_ch(O, rec);
        const float vm = D ? filter_ray(O, *D, chm, rec) : filter_point(O, chm);
        if (vm < 0.f) continue;
        const uint32_t i = sc.midx[c * RT_CLUSTER + m];
        if (i < sc.n) fn(i);
      }
    }
  };
  if (s.kind == K_TRACE) {
    DirQ D;
    s.minT = 1000.f; s.hitIdx = -1;
    if (!make_dir(D, s.qo, s.rayD)) return;
    auto test = [&](uint32_t i) {
      ctr.exactTests++;
      float t;
      if (ray_sphere_exact(sc.geo[i], s.qo, s.rayD, t) && (t < s.minT || (t == s.minT && (int)i < s.hitIdx))) { s.minT = t; s.hitIdx = (int)i; }
    };
    if (ofil && dir_filterable(D)) {
      culled(&D, test);
      if (audit) for (uint32_t i = 0; i < sc.n; ++i) { float t; if (ray_sphere_exact(sc.geo[i], s.qo, s.rayD, t) && !flagged[clusterOf[i]]) ++violations; }
    } else brute(test);
  } else if (s.kind == K_SHADOW) {
    s.blocked = 0u;
    shadow_geo(s, sc, sg);
    for (int k = 0; k < s.ndirs; ++k) {
      DirQ D;
      if (!make_dir(D, org, sg.d[k])) continue;
      auto test = [&](uint32_t i) {
        if ((s.blocked >> k) & 1u) return;
        ctr.exactTests++;
        if (resolve_shadow(org, sg.d[k], sg.gap[k], sc.geo[i])) s.blocked |= 1u << k;
      };
      if (ofil && dir_filterable(D)) {
        culled(&D, test);
        if (audit) for (uint32_t i = 0; i < sc.n; ++i) { float t; if (ray_sphere_exact(sc.geo[i], org, sg.d[k], t) && !flagged[clusterOf[i]]) ++violations; }
      } else brute(test);
    }
  } else if (s.kind == K_CONTAIN) {
    s.hitIdx = -1;
    auto test = [&](uint32_t i) {
      ctr.exactTests++;
      if (contains_exact(sc.geo[i], s.qo) && (s.hitIdx < 0 || (int)i < s.hitIdx)) s.hitIdx = (int)i;
    };
    if (ofil) {
      culled(nullptr, test);
      if (audit) for (uint32_t i = 0; i < sc.n; ++i) if (contains_exact(sc.geo[i], s.qo) && !flagged[clusterOf[i]]) ++violations;
    } else brute(test);
  }
}

/* Answer the slot's pending query the way one pass of the kernel does. */
static void answer(Slot& s, const SceneView& sc, Counters& ctr, bool noFilter, ShadowGeo& sg) {
  const V3 org = (s.kind == K_SHADOW) ? s.P : s.qo;
  const OriginQ O = make_origin(org);
  const bool ofil = origin_filterable(O);
  if (s.kind == K_TRACE) {
    DirQ D;
    s.minT = 1000.f; s.hitIdx = -1;
    if (!make_dir(D, s.qo, s.rayD)) return;           /* zero direction: certain miss */
    const bool fil = !noFilter && ofil && dir_filterable(D);
    for (uint32_t i = 0; i < (fil ? sc.nPad : sc.n); ++i) {
      if (fil && filter_ray(O, D, filter_ch(O, sc.filt[i]), sc.filt[i]) < 0.f) continue;
      if (i >= sc.n) continue;
      ctr.exactTests++;
      resolve_trace(s.minT, s.hitIdx, s.qo, s.rayD, sc.geo[i], i);
    }
  } else if (s.kind == K_SHADOW) {
    s.blocked = 0u;
    shadow_geo(s, sc, sg);
    for (int k = 0; k < s.ndirs; ++k) {
      DirQ D;
      if (!make_dir(D, org, sg.d[k])) continue;
      const bool fil = !noFilter && ofil && dir_filterable(D);
      for (uint32_t i = 0; i < (fil ? sc.nPad : sc.n); ++i) {
        if (fil && filter_ray(O, D, filter_ch(O, sc.filt[i]), sc.filt[i]) < 0.f) continue;
        if (i >= sc.n) continue;
        if ((s.blocked >> k) & 1u) break;
        ctr.exactTests++;
        if (resolve_shadow(org, sg.d[k], sg.gap[k], sc.geo[i])) s.blocked |= 1u << k;
      }
    }
  } else if (s.kind == K_CONTAIN) {
    s.hitIdx = -1;
    const bool fil = !noFilter && ofil;
    for (uint32_t i = 0; i < (fil ? sc.nPad : sc.n); ++i) {
      if (fil && filter_point(O, filter_ch(O, sc.filt[i])) < 0.f) continue;
      if (i >= sc.n) continue;
      ctr.exactTests++;
      resolve_contain(s.hitIdx, s.qo, sc.geo[i], i);
    }
  }
}

extern "C" int hostsim_render(const rt_sphere* spheres, unsigned n, const rt_light* lights,
                              unsigned nl, unsigned W, unsigned H, float zoom, float alias, int S,
                              unsigned row_begin, unsigned row_count, unsigned row_step,
                              float* out, uint64_t* counters /* [10] or NULL */, int mode /* 0 filter, 1 exact only, 2 accelerated, 3 accelerated + audit */) {
  if (!out || W == 0 || H == 0 || S < 1 || S > RT_MAX_STACK) return -1;
  if (row_step == 0) row_step = 1;
  std::vector<float4_> h;
  SceneLayout lay;
  build_scene_soa(spheres, n, lights, nl, h, lay);
  const SceneView sc = scene_view(h.data(), lay);
  const Camera cam = make_camera(W, H, zoom, alias, S, (int)sc.n);
  uint64_t c[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<uint32_t> clusterOf(sc.n, 0u);
  for (uint32_t k = 0; k < sc.nc * RT_CLUSTER; ++k)
    if (sc.midx[k] < sc.n) clusterOf[sc.midx[k]] = k / RT_CLUSTER;

#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c[:10])
  for (long k = 0; k < (long)row_count; ++k) {
    const unsigned gy = row_begin + (unsigned)k * row_step;
    for (unsigned gx = 0; gx < W; ++gx) {
      Slot s;
      Frame stack[RT_MAX_STACK];
      Counters ctr; memset(&ctr, 0, sizeof ctr);
      float* px = out + ((size_t)k * W + gx) * 3;
      V3 samples[64 * 64];
      int count = 0;
      for (int si = 0; si < cam.nIter && si < 64; ++si) {
        for (int sj = 0; sj < cam.nIter && sj < 64; ++sj) {
          memset(&s, 0, sizeof s);
          start_task(s, ctr, cam, gx, gy, 0, si, sj);
          for (;;) {
            c[6]++; c[7]++;
            ShadowGeo sg;
            if (mode >= 2) answer_accel(s, sc, ctr, sg, mode == 3, clusterOf, c[8], c[9]);
            else answer(s, sc, ctr, mode == 1, sg);
            if (advance(s, stack, ctr, sc, cam, &sg)) break;
            /* between two passes the kernel keeps only the 21-word record (slot_pack / slot_unpack):
             * everything the record drops must really be dead */
            uint32_t rec[RT_SLOT_WORDS];
            slot_pack<1>(rec, s);
            memset(&s, 0xA5, sizeof s);
            slot_unpack<1>(rec, s);
          }
          samples[count++] = sample_value(s, cam);
        }
      }
      const V3 acc = combine_samples(samples, count);
      px[0] = acc.x; px[1] = acc.y; px[2] = acc.z;
      c[0] += ctr.rays; c[1] += ctr.shadow; c[2] += ctr.containQ; c[3] += ctr.containT;
      c[4] += ctr.exactTests; c[5] += ctr.samples;
    }
  }
  if (counters) memcpy(counters, c, sizeof c);
  return 0;
}

/* The cluster form of a scene (rt_soa.h build_clusters), for tests/test_hostsim.py:
 * out_rec[nc*4] = the clusters' {Cx, Cy, Cz, w}, out_idx[nc*RT_CLUSTER] = member sphere indices
 * (0x3FFF = padding).  Returns the number of clusters, or -1. */
extern "C" int hostsim_clusters(const rt_sphere* spheres, unsigned n, float* out_rec, unsigned short* out_idx,
                                unsigned capacity) {
  std::vector<float4_> h;
  SceneLayout lay;
  build_scene_soa(spheres, n, nullptr, 0, h, lay);
  const SceneView sc = scene_view(h.data(), lay);
  if (sc.nc > capacity) return -1;
  for (uint32_t c = 0; c < sc.nc; ++c) {
    out_rec[4 * c + 0] = sc.cfilt[c].x; out_rec[4 * c + 1] = sc.cfilt[c].y;
    out_rec[4 * c + 2] = sc.cfilt[c].z; out_rec[4 * c + 3] = sc.cfilt[c].w;
    for (uint32_t m = 0; m < RT_CLUSTER; ++m) out_idx[c * RT_CLUSTER + m] = sc.midx[c * RT_CLUSTER + m];
  }
  for (uint32_t c = sc.nc; c < sc.ncPad; ++c)            /* padding clusters must never flag */
    if (!(sc.cfilt[c].w == INFINITY)) return -2;
  return (int)sc.nc;
}

/* The work map of rt_core.cuh (groups of 32 items per tile; 8x4 tiles x samples, or 16x8 tiles x four pixel
 * sub-lattices at 1 spp): every result record must be produced by exactly one item, and tile_of_dst must
 * invert the map.  hits[localRows*W*spp] receives the number of items per record; returns the number of
 * items whose tile_of_dst disagrees or whose frame row is not one of this shard's rows (must be 0). */
extern "C" int hostsim_workmap(unsigned W, unsigned H, unsigned stripRows, unsigned stripFirst, unsigned stripStride,
                               unsigned spp, unsigned nIter, unsigned* hits, unsigned* outLocalRows) {
  const unsigned nStrips = (H + stripRows - 1) / stripRows;
  unsigned localRows = 0;
  for (unsigned s = stripFirst; s < nStrips; s += stripStride) localRows += (s * stripRows + stripRows <= H) ? stripRows : H - s * stripRows;
  *outLocalRows = localRows;
  const WorkMap m = make_workmap(W, localRows, stripRows, stripFirst, stripStride, spp, nIter);
  int bad = 0;
  for (uint32_t t = 0; t < m.nTiles; ++t)
    for (uint32_t k = 0; k < m.K; ++k)
      for (uint32_t wi = 0; wi < 32; ++wi) {
        uint32_t gx, gy, dst; int si, sj;
        if (!work_item(m, t, k, wi, gx, gy, dst, si, sj)) continue;
        if (hits) hits[dst]++;
        if (tile_of_dst(m, dst) != t) ++bad;
        if (gx >= W || gy >= H || (gy / stripRows) % stripStride != stripFirst) ++bad;
        if ((unsigned)(si * (int)nIter + sj) != (m.lattice ? 0u : k)) ++bad;
      }
  return bad;
}
