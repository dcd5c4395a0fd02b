/* rt_wavefront.cuh — the wavefront engine of the trace loop (large frames).
 *
 * The persistent kernel (rt_kernels.cuh: trace_kernel) keeps every sample's state next to the
 * lane that traces it and lets warps vote for a query kind.  That leaves three sources of
 * idle FP32 issue: lanes whose slots wait on another kind, warps stuck in the serial O(1)
 * shading code while they hold registers sized for it, and only 16 warps per SM.  For large
 * frames this engine separates the two kinds of work into different kernels:
 *
 *   wf_filter<KIND>   PURE sphere passes.  The pending queries of a pool of samples sit in
 *                     per-kind queues, so every lane of every warp carries a query of the
 *                     same kind (trace: two rays per lane, shadow: one batch of <= 4 rays,
 *                     contain: two probes).  Same packed FFMA2 loops, gather and exact
 *                     resolve as the persistent kernel (pass_trace / pass_shadow /
 *                     pass_contain are reused as they are); the answers go back to the pool.
 *   wf_shade          one thread per answered sample: the O(1) state transition of
 *                     rt_core.cuh (advance), sample completion + refill from the frame's
 *                     work counter, and the push of the next query into its kind's queue
 *                     (warp-aggregated atomics).
 *
 * The pool (samples in flight) lives in global memory as a structure of arrays — word w of
 * sample i at st[w * P + i] — so both kernels access it fully coalesced; call stacks are one
 * contiguous 1 KB record per sample.  The host loop (rt_shim.cu) alternates the kernels
 * until the queues are empty.  Per-sample arithmetic is the same code as everywhere else, so
 * the framebuffer is bit-identical to the persistent kernel's and to the oracle's.
 */
#ifndef RT_WAVEFRONT_CUH
#define RT_WAVEFRONT_CUH

#include <stddef.h>
#include "rt_kernels.cuh"

namespace rtg {

#define RT_SLOT_WORDS ((int)(sizeof(Slot) / sizeof(uint32_t)))
#define RT_W(field) ((int)(offsetof(Slot, field) / sizeof(uint32_t)))

struct WfParams {
  SceneView sc;
  Camera cam;
  uint32_t* st;              /* [RT_SLOT_WORDS][P] pool, structure of arrays           */
  Frame* stacks;             /* [P][RT_MAX_STACK] suspended calls                       */
  uint32_t* queues;          /* [2][3][P] sample ids waiting on trace / shadow / contain */
  uint32_t* counts;          /* [2][4] queue lengths (index kind-1)                     */
  uint32_t P;                /* pool size                                               */
  uint32_t cur;              /* buffer holding the CURRENT queries; shade fills 1-cur   */
  float4* fb;
  float4* samples;
  uint32_t spp;
  unsigned int* workCounter; /* next work item (sample) of the frame                    */
  unsigned int* maxBits;
  unsigned long long* counters;
  uint32_t localRows, stripRows, stripFirst, stripStride, totalWork;
  int noFilter;
};

__device__ __forceinline__ uint32_t* wf_queue(const WfParams& p, uint32_t buf, int kind) {
  return p.queues + ((size_t)buf * 3u + (uint32_t)(kind - 1)) * p.P;
}
__device__ __forceinline__ uint32_t wf_ld(const WfParams& p, int word, uint32_t id) { return p.st[(size_t)word * p.P + id]; }
__device__ __forceinline__ float wf_ldf(const WfParams& p, int word, uint32_t id) { return __uint_as_float(wf_ld(p, word, id)); }
__device__ __forceinline__ void wf_st(const WfParams& p, int word, uint32_t id, uint32_t v) { p.st[(size_t)word * p.P + id] = v; }
__device__ __forceinline__ V3 wf_ldv(const WfParams& p, int word, uint32_t id) {
  return mk(wf_ldf(p, word, id), wf_ldf(p, word + 1, id), wf_ldf(p, word + 2, id));
}
__device__ __forceinline__ void wf_slot_load(const WfParams& p, Slot& s, uint32_t id) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&s);
#pragma unroll
  for (int i = 0; i < RT_SLOT_WORDS; ++i) w[i] = wf_ld(p, i, id);
}
__device__ __forceinline__ void wf_slot_store(const WfParams& p, const Slot& s, uint32_t id) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&s);
#pragma unroll
  for (int i = 0; i < RT_SLOT_WORDS; ++i) wf_st(p, i, id, w[i]);
}

/* Work item -> sample.  Row-major over the shard's pixels, samples of a pixel adjacent
 * (no tiling needed here: ray coherence does not matter to the filter kernels). */
__device__ __forceinline__ void wf_task(const WfParams& p, uint32_t idx, uint32_t& gx, uint32_t& gy,
                                        uint32_t& dst, int& si, int& sj) {
  const uint32_t px = idx / p.spp, k = idx - px * p.spp;
  const uint32_t y = px / p.cam.W, x = px - y * p.cam.W;
  const uint32_t strip = y / p.stripRows;
  gx = x;
  gy = (strip * p.stripStride + p.stripFirst) * p.stripRows + (y - strip * p.stripRows);
  dst = idx;                                   /* = (y * W + x) * spp + k */
  si = (int)(k / (uint32_t)p.cam.nIter);
  sj = (int)(k - (uint32_t)si * (uint32_t)p.cam.nIter);
}

/* Per-thread tallies -> the frame's counters (same slots as trace_kernel's). */
__device__ __forceinline__ void wf_flush_counters(const WfParams& p, const Counters& ctr, float laneMax) {
  unsigned mb = __reduce_max_sync(RT_FULL, __float_as_uint(laneMax));
  unsigned long long v[7];
  v[0] = ctr.rays; v[1] = ctr.shadow; v[2] = ctr.containQ; v[3] = ctr.containT;
  v[4] = ctr.exactTests; v[5] = ctr.samples; v[6] = ctr.nullRays;
#pragma unroll
  for (int i = 0; i < 7; ++i)
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(RT_FULL, v[i], o);
  if ((threadIdx.x & 31u) == 0) {
    if (mb) atomicMax(p.maxBits, mb);
#pragma unroll
    for (int i = 0; i < 7; ++i)
      if (v[i]) atomicAdd(&p.counters[i], v[i]);
  }
}

/* ---- first wave: P samples take the first P work items ------------------------------- */
__global__ void __launch_bounds__(256) wf_spawn(const WfParams p) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  Counters ctr;
  ctr.rays = ctr.shadow = ctr.containQ = ctr.containT = ctr.exactTests = ctr.samples = ctr.nullRays = 0;
  if (id < p.P) {
    Slot s;
    uint32_t* w = reinterpret_cast<uint32_t*>(&s);
#pragma unroll
    for (int i = 0; i < RT_SLOT_WORDS; ++i) w[i] = 0u;
    uint32_t gx, gy, dst;
    int si, sj;
    wf_task(p, id, gx, gy, dst, si, sj);
    start_task(s, ctr, p.cam, gx, gy, dst, si, sj);
    wf_slot_store(p, s, id);
    wf_queue(p, p.cur, K_TRACE)[id] = id;      /* counts[cur][trace] = P is set by the host */
  }
  wf_flush_counters(p, ctr, 0.f);
}

/* ---- filter kernels ------------------------------------------------------------------ */
/* Shared-memory layout and scene staging are the persistent kernel's. */
template <int KIND>
__global__ void __launch_bounds__(RT_BLOCK, 3) wf_filter(const WfParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t count = p.counts[p.cur * 4u + (uint32_t)(KIND - 1)];
  constexpr uint32_t PER = (KIND == K_SHADOW) ? 1u : 2u;       /* queries per lane */
  const uint32_t perBlock = RT_BLOCK * PER;
  if ((size_t)blockIdx.x * perBlock >= count) return;          /* nothing for this CTA */

  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float4* sFilt = reinterpret_cast<float4*>(smem_raw + 16);
  const uint32_t filtBytes = p.sc.nPad * 16u;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && filtBytes) {
    mbar_expect_tx(bar, filtBytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.sc.filt);
    unsigned char* dstp = reinterpret_cast<unsigned char*>(sFilt);
    for (uint32_t off = 0; off < filtBytes; off += 32768u) {
      const uint32_t n = (filtBytes - off < 32768u) ? (filtBytes - off) : 32768u;
      tma_bulk_g2s(dstp + off, src + off, n, bar);
    }
  }
  if (filtBytes) mbar_wait(bar, 0);

  WarpCtx w;
  w.filt = sFilt;
  w.list = reinterpret_cast<unsigned short*>(smem_raw + 16 + filtBytes);
  w.geo = reinterpret_cast<float*>(smem_raw + 16 + filtBytes + RT_LIST_MAX * RT_BLOCK * sizeof(unsigned short));
  w.tid = tid;
  w.nPad = p.sc.nPad;
#ifdef RT_PHASE_TIMING
  w.t0 = clock64();
  for (int i = 0; i < 6; ++i) w.phase[i] = 0;
#endif
  TraceParams tp;
  tp.sc = p.sc; tp.cam = p.cam; tp.noFilter = p.noFilter;

  Counters ctr;
  ctr.rays = ctr.shadow = ctr.containQ = ctr.containT = ctr.exactTests = ctr.samples = ctr.nullRays = 0;
  const uint32_t* q = wf_queue(p, p.cur, KIND);

  for (uint32_t base = blockIdx.x * perBlock; base < count; base += gridDim.x * perBlock) {
    Slot sl[2];
    const uint32_t i0 = base + tid * PER;
    const bool have0 = i0 < count, have1 = (PER == 2u) && (i0 + 1u < count);
    const uint32_t id0 = have0 ? q[i0] : 0u, id1 = have1 ? q[i0 + 1u] : 0u;
    if (KIND == K_TRACE) {
      if (have0) { sl[0].kind = K_TRACE; sl[0].ndirs = 1; sl[0].qo = wf_ldv(p, RT_W(qo), id0); sl[0].rayD = wf_ldv(p, RT_W(rayD), id0); }
      if (have1) { sl[1].kind = K_TRACE; sl[1].ndirs = 1; sl[1].qo = wf_ldv(p, RT_W(qo), id1); sl[1].rayD = wf_ldv(p, RT_W(rayD), id1); }
      pass_trace<false>(tp, w, sl, have0 ? 0 : -1, have1 ? 1 : -1, ctr);
      if (have0) { wf_st(p, RT_W(minT), id0, __float_as_uint(sl[0].minT)); wf_st(p, RT_W(hitIdx), id0, (uint32_t)sl[0].hitIdx); }
      if (have1) { wf_st(p, RT_W(minT), id1, __float_as_uint(sl[1].minT)); wf_st(p, RT_W(hitIdx), id1, (uint32_t)sl[1].hitIdx); }
    } else if (KIND == K_SHADOW) {
      int nd = 0;
      if (have0) {
        sl[0].kind = K_SHADOW; sl[0].P = wf_ldv(p, RT_W(P), id0);
        sl[0].light = (int)wf_ld(p, RT_W(light), id0); sl[0].ndirs = (int)wf_ld(p, RT_W(ndirs), id0);
        nd = sl[0].ndirs;
      }
      const int ndMax = __reduce_max_sync(RT_FULL, nd);
      if (ndMax <= 2) pass_shadow<false, 2>(tp, w, sl, have0 ? 0 : -1, ctr);
      else            pass_shadow<false, 4>(tp, w, sl, have0 ? 0 : -1, ctr);
      if (have0) wf_st(p, RT_W(blocked), id0, sl[0].blocked);
    } else {
      if (have0) { sl[0].kind = K_CONTAIN; sl[0].qo = wf_ldv(p, RT_W(qo), id0); }
      if (have1) { sl[1].kind = K_CONTAIN; sl[1].qo = wf_ldv(p, RT_W(qo), id1); }
      pass_contain<false>(tp, w, sl, have0 ? 0 : -1, have1 ? 1 : -1, ctr);
      if (have0) wf_st(p, RT_W(hitIdx), id0, (uint32_t)sl[0].hitIdx);
      if (have1) wf_st(p, RT_W(hitIdx), id1, (uint32_t)sl[1].hitIdx);
    }
  }
  wf_flush_counters(p, ctr, 0.f);
}

/* ---- shade kernel --------------------------------------------------------------------- */
/* One thread per sample whose query of kind `kind` was just answered.  All lanes of a warp
 * advance the same kind, so the paths agree as far as the data lets them. */
__global__ void __launch_bounds__(128) wf_shade(const WfParams p, const int kind) {
  const uint32_t count = p.counts[p.cur * 4u + (uint32_t)(kind - 1)];
  const uint32_t* q = wf_queue(p, p.cur, kind);
  const uint32_t nxt = p.cur ^ 1u;
  Counters ctr;
  ctr.rays = ctr.shadow = ctr.containQ = ctr.containT = ctr.exactTests = ctr.samples = ctr.nullRays = 0;
  float laneMax = 0.f;
  const uint32_t lane = threadIdx.x & 31u;

  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const uint32_t id = q[i];
    Slot s;
    wf_slot_load(p, s, id);
    Frame* stack = p.stacks + (size_t)id * RT_MAX_STACK;
    ShadowGeo sg;
    if (s.kind == K_SHADOW) shadow_geo(s, p.sc, sg);
    bool alive = true;
    if (advance(s, stack, ctr, p.sc, p.cam, &sg)) {
      /* sample finished: record it, then take the next work item of the frame, if any */
      const V3 v = sample_value(s, p.cam);
      if (p.spp == 1u) {
        const V3 a = vadd(mk(0.f, 0.f, 0.f), v);           /* pixel = 0 + sample (main.cpp:420,446) */
        p.fb[s.pixel] = make_float4(a.x, a.y, a.z, 1.f);
        if (a.x > laneMax) laneMax = a.x;                   /* algebra.h:74-82, NaN skipped */
        if (a.y > laneMax) laneMax = a.y;
        if (a.z > laneMax) laneMax = a.z;
      } else {
        p.samples[s.pixel] = make_float4(v.x, v.y, v.z, 1.f);
      }
      const unsigned m = __activemask();
      const int leader = __ffs(m) - 1;
      uint32_t b = 0;
      if ((int)lane == leader) b = atomicAdd(p.workCounter, (unsigned)__popc(m));
      b = __shfl_sync(m, b, leader);
      const uint32_t wi = b + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (wi < p.totalWork) {
        uint32_t gx, gy, dst;
        int si, sj;
        wf_task(p, wi, gx, gy, dst, si, sj);
        start_task(s, ctr, p.cam, gx, gy, dst, si, sj);
      } else {
        alive = false;
        s.kind = K_NULL;
      }
    }
    if (alive) {
      /* push the sample into the queue of its next query, one atomic per kind per warp */
      const unsigned m = __activemask();
      const unsigned peers = __match_any_sync(m, s.kind);
      const int leader = __ffs(peers) - 1;
      uint32_t b = 0;
      if ((int)lane == leader) b = atomicAdd(&p.counts[nxt * 4u + (uint32_t)(s.kind - 1)], (unsigned)__popc(peers));
      b = __shfl_sync(peers, b, leader);
      wf_queue(p, nxt, s.kind)[b + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = id;
      wf_slot_store(p, s, id);
    }
  }
  wf_flush_counters(p, ctr, laneMax);
}

}  // namespace rtg
#endif
