/* rt_kernels.cuh — sm_100a kernels of the trace loop.
 *
 * trace_kernel     persistent CTAs; every lane owns NSLOTS samples in flight.  Replaces
 *                  `__kernel raytrace` (raytrace_kernel.cl:870-973) and the CPU pixel
 *                  loop (main.cpp:404-453).  ACCEL = the optional two-level cluster filter.
 * combine_kernel   sums each pixel's samples in the reference's order, takes the frame maximum
 * pack_kernel      float4 framebuffer -> packed Vec[W*H] (the reference's dst layout)
 * quantise_kernel  float4 framebuffer + max -> RGB8 (main.cpp:71-76)
 * assemble_rgb8_kernel  multi-GPU strip de-interleave
 *
 * One pass of trace_kernel (per warp, all lanes converged throughout):
 *   refill   free slots take the next samples from a tile queue (one global atomicAdd
 *            per warp granule, __ballot_sync ranks the takers)
 *   vote     every lane reports which query kinds its slots are waiting on; the warp
 *            picks the kind that fills most lanes (__reduce_add_sync)
 *   filter   the chosen kind's loop over ALL spheres, sphere records staged once per
 *            CTA into shared memory by a TMA bulk copy (cp.async.bulk + mbarrier) or
 *            read from __constant__ for small scenes:
 *              trace    2 rays per lane    (1 LDS.128 + 7 FFMA2 + FADD2 + 2 SHF) per sphere
 *              shadow   4 rays, one origin (1 LDS.128 + 3 FFMA + FADD + 2 x [4 FFMA2 + 2 SHF])
 *              contain  2 probes per lane  (1 LDS.128 + 3 FFMA2 + FADD2 + 2 SHF)
 *            (packed FP32 pairs: one FFMA2 serves both rays) — each test leaves one SIGN BIT
 *            (certain miss or not) in the group's funnel-shifted register
 *   gather   set bits become (sub-query, sphere) entries in a per-lane shared-memory list
 *   resolve  k-th entries of all lanes go through the reference's exact expressions together
 *   advance  the served slots take their O(1) shading / state transition (one kind per
 *            pass, so the lanes agree on the path)
 */
#ifndef RT_KERNELS_CUH
#define RT_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "rt_core.cuh"

namespace rtg {

#ifndef RT_BLOCK
#define RT_BLOCK 256
#endif
#define RT_LIST_MAX 24
/* accelerated mode: its loops run over an eighth of the records, so smaller unrolled groups and a single
 * shadow instance (fewer instructions to fetch per pass) win: 41.9 -> 38.8 ms at 4K / 1 024 spheres */
#ifndef RT_ACCEL_ONE_SHADOW
#define RT_ACCEL_ONE_SHADOW 1
#endif
#ifndef RT_GROUP_TA
#define RT_GROUP_TA 8
#endif
#ifndef RT_GROUP_CA
#define RT_GROUP_CA 8
#endif
#ifndef RT_GROUP_S4A
#define RT_GROUP_S4A 8
#endif
/* accelerated mode: (sub, cluster) entries per lane between two flushes (TraceParams.list1Max, >= 32 because one
 * group can add 32): 40 is best at 1 024 spheres (more L1 left for the slots), 64 at 4 096 (-10 %, fewer flushes) */
#define RT_LIST1_SMALL 40
#define RT_LIST1_LARGE 64
#define RT_CONST_MAX_SPHERES 1024
#define RT_NUM_COUNTERS 24
#define RT_NO_PIXEL 0xFFFFFFFFu

struct TraceParams {
  SceneView sc;
  Camera cam;
  float4* fb;               /* [localRows*W] {r,g,b,1}                              */
  float4* samples;          /* [localRows*W*spp] scaled sample values (spp > 1 only) */
  uint32_t spp;             /* samples per pixel = nIter * nIter                    */
  unsigned int* workCounter;/* tile queue head                                      */
  unsigned int* maxBits;    /* running max of positive channel values (float bits)  */
  unsigned long long* counters;  /* [RT_NUM_COUNTERS], see rt_shim.cu                */
  uint32_t localRows;       /* rows rendered by this context                        */
  uint32_t stripRows, stripFirst, stripStride;   /* row r is ours iff (r/stripRows)%stripStride==stripFirst */
  uint32_t tilesX, totalWork, chunk;
  int noFilter;             /* debug: exact test for every sphere                   */
  int prefetch;             /* prefetch served slots' state into L1 before the sphere loop */
  uint32_t list1Max;        /* accelerated mode: capacity of the per-lane (sub, cluster) lists */
};

/* "__constant__ staging" (option staging=1): the filter records travel with the launch as a
 * __grid_constant__ kernel parameter, i.e. in the constant bank the parameters live in (LDC with a
 * warp-uniform index), so every launch reads ITS context's records — a module-global __constant__
 * symbol would be shared by all contexts of a process. */
struct ConstRecords { float4_ r[RT_CONST_MAX_SPHERES]; };

/* ---- TMA bulk staging (global -> shared, completion on an mbarrier) --------- */
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

/* Map a queue index to a work item = one sample of one pixel.  Pixels are walked in 8x4
 * tiles (row-major inside, tiles row-major); 32 consecutive indices are the 32 pixels of a
 * tile for one sample, the next 32 the same tile's next sample. */
__device__ __forceinline__ bool work_to_task(const TraceParams& p, uint32_t idx, uint32_t& gx,
                                             uint32_t& gy, uint32_t& dst, int& si, int& sj) {
  const uint32_t grp = idx >> 5, within = idx & 31u;
  const uint32_t tile = grp / p.spp, k = grp - tile * p.spp;
  const uint32_t ty = tile / p.tilesX, tx = tile - ty * p.tilesX;
  const uint32_t x = tx * 8u + (within & 7u);
  const uint32_t y = ty * 4u + (within >> 3);
  if (x >= p.cam.W || y >= p.localRows) return false;
  const uint32_t strip = y / p.stripRows;
  gx = x;
  gy = (strip * p.stripStride + p.stripFirst) * p.stripRows + (y - strip * p.stripRows);
  dst = (y * p.cam.W + x) * p.spp + k;
  si = (int)(k / (uint32_t)p.cam.nIter);
  sj = (int)(k - (uint32_t)si * (uint32_t)p.cam.nIter);
  return true;
}

/* Per-warp state shared by the passes. */
struct WarpCtx {
#ifdef RT_PHASE_TIMING
  long long t0;
  long long phase[6];       /* tail (cycles after the queue ran dry), set-up, filter loop, resolve, advance, longest tail */
  long long tDry;
#endif
  const float4* filt;       /* filter records (shared memory, or unused with __constant__); accelerated mode: cluster records */
  const float4* mfilt;      /* accelerated mode: the clusters' member records (shared memory) */
  const unsigned short* midx;   /* accelerated mode: the members' sphere indices (shared memory) */
  unsigned short* list1;    /* accelerated mode: per-lane (sub, cluster) lists: list1[k * RT_BLOCK + tid] */
  unsigned short* list;     /* per-lane candidate lists: list[k * RT_BLOCK + tid]          */
  float* geo;               /* per-lane shadow-batch rays of the current pass: geo[w * RT_BLOCK + tid], 16 words */
  uint32_t tid;
  uint32_t nPad;
};

template <bool USE_CONST>
__device__ __forceinline__ float4_ load_filt(const WarpCtx& w, const ConstRecords& cr, uint32_t i) {
  float4_ s;
  if (USE_CONST) {
    s = cr.r[i];          /* cr IS the kernel's __grid_constant__ parameter: LDC with a warp-uniform index */
  } else {
    const float4 v = w.filt[i];
    s.x = v.x; s.y = v.y; s.z = v.z; s.w = v.w;
  }
  return s;
}

/* ---- packed FP32 pairs (Blackwell FFMA2 / FADD2) --------------------------------------
 * sm_100a executes fma.rn.f32x2 as ONE instruction (SASS FFMA2) on a register pair, with a
 * broadcast form for scalar operands.  The filter loops keep two rays per lane in the two
 * halves, so every multiply-add of the discriminant is issued once for both: the loops
 * stop being issue-bound and run against the FP32 pipe itself. */
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 pk1(float x) { return pk(x, x); }
__device__ __forceinline__ float lo_of(f32x2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi_of(f32x2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
struct Origin2 { f32x2 px, py, pz, nq; };
struct Dir2 { f32x2 ndx, ndy, ndz, od; };
__device__ __forceinline__ Origin2 pack_origin(const OriginQ& a, const OriginQ& b) {
  Origin2 o; o.px = pk(a.px, b.px); o.py = pk(a.py, b.py); o.pz = pk(a.pz, b.pz); o.nq = pk(a.nq, b.nq); return o;
}
__device__ __forceinline__ Dir2 pack_dir(const DirQ& a, const DirQ& b) {
  Dir2 d; d.ndx = pk(a.ndx, b.ndx); d.ndy = pk(a.ndy, b.ndy); d.ndz = pk(a.ndz, b.ndz); d.od = pk(a.od, b.od); return d;
}
/* b' = d'.(o - c) for two rays */
__device__ __forceinline__ f32x2 bq2(const Dir2& D, f32x2 cx, f32x2 cy, f32x2 cz) {
  f32x2 b = fma2(D.ndx, cx, D.od);
  b = fma2(D.ndy, cy, b);
  return fma2(D.ndz, cz, b);
}

/* Spheres per unrolled group of each pass.  Small groups keep the three hot loops inside
 * the instruction caches (they run concurrently on one SM in different warps). */
#ifndef RT_GROUP_T
#define RT_GROUP_T 16
#endif
#ifndef RT_GROUP_S4
#define RT_GROUP_S4 8
#endif
#ifndef RT_GROUP_S2
#define RT_GROUP_S2 16
#endif
#ifndef RT_GROUP_C
#define RT_GROUP_C 16
#endif

/* One group's sign bits are collected in ONE register: the loop shifts in, sphere by sphere,
 * the sign of each of the lane's ND sub-queries (funnel shift, one instruction per test;
 * every instruction that is not an FMA costs the FMA pipe an issue cycle, scripts/ubench_mix.cu).
 * After the group, test (j, sub) sits at bit G*ND-1 - (j*ND + sub).  mask_of_sub gives the bits
 * of one sub-query; gather turns the set bits of ~signs & mask into list entries
 * (sub << 14 | sphere): scanning from the top bit yields each sub-query's spheres in
 * increasing order.  A full list sets `overflow`: that lane then resolves the pass exactly
 * against every sphere (rare). */
template <int ND, int G>
__host__ __device__ constexpr unsigned mask_of_sub(int sub) {
  unsigned m = 0u;
  for (int j = 0; j < G; ++j) m |= 1u << (G * ND - 1 - (j * ND + sub));
  return m;
}
/* keeps a loop-invariant value in its register (the compiler otherwise rebuilds it from
 * predicates / the constant bank on every trip, which costs issue slots in the hot loops) */
__device__ __forceinline__ unsigned pin(unsigned v) { asm volatile("" : "+r"(v)); return v; }

template <int ND, int G>
__device__ __forceinline__ void gather(const WarpCtx& w, unsigned comb, uint32_t base, int& cnt,
                                       bool& overflow) {
  static_assert(G * ND <= 32 && (ND & (ND - 1)) == 0, "one 32-bit register per group");
  while (comb) {
    const int b = 31 - __clz(comb);
    comb &= ~(1u << b);
    if (cnt < RT_LIST_MAX) {
      const uint32_t idx = (uint32_t)(G * ND - 1 - b);
      const uint32_t sub = idx % ND, j = idx / ND;
      w.list[cnt * RT_BLOCK + w.tid] = (unsigned short)((sub << 14) | (base + j));
      ++cnt;
    } else {
      overflow = true;
    }
  }
}

#define RT_FULL 0xFFFFFFFFu

/* Optional phase timing (development builds, -DRT_PHASE_TIMING): wall cycles a warp spends in
 * each phase of a pass, accumulated into counters[16..21]. */
#ifdef RT_PHASE_TIMING
#define RT_TICK(slot) do { const long long now_ = clock64(); w.phase[slot] += now_ - w.t0; w.t0 = now_; } while (0)
#else
#define RT_TICK(slot) do { } while (0)
#endif

/* Cold path (list overflow, non-finite geometry, no_filter debug mode): answer the slot's
 * pending query with the exact test against every sphere.  Out of line; returns the
 * number of exact tests. */
__device__ __noinline__ uint32_t exact_all(const SceneView sc, Slot* s, unsigned subs) {
  uint32_t tests = 0;
  if (s->kind == K_TRACE) {
    float t = 1000.f; int h = -1;
    DirQ D;
    if (make_dir(D, s->qo, s->rayD)) {
      for (uint32_t i = 0; i < sc.n; ++i) resolve_trace(t, h, s->qo, s->rayD, sc.geo[i], i);
      tests = sc.n;
    }
    s->minT = t; s->hitIdx = h;
  } else if (s->kind == K_SHADOW) {
    unsigned blocked = s->blocked;
    ShadowGeo sg;
    shadow_geo(*s, sc, sg);
    for (int k = 0; k < RT_SHADOW_BATCH; ++k) {
      if ((subs >> k) & 1u) {
        blocked &= ~(1u << k);
        for (uint32_t i = 0; i < sc.n; ++i) {
          ++tests;
          if (resolve_shadow(s->P, sg.d[k], sg.gap[k], sc.geo[i])) { blocked |= 1u << k; break; }
        }
      }
    }
    s->blocked = blocked;
  } else if (s->kind == K_CONTAIN) {
    int h = -1;
    for (uint32_t i = 0; i < sc.n && h < 0; ++i) { ++tests; resolve_contain(h, s->qo, sc.geo[i], i); }
    s->hitIdx = h;
  }
  return tests;
}

/* ---- trace pass: up to two rays per lane ---------------------------------------- */
template <bool USE_CONST>
__device__ __forceinline__ void pass_trace(const TraceParams& p, const ConstRecords& cr, WarpCtx& w, Slot* slots, int s0,
                                           int s1, Counters& ctr) {
  constexpr int G = RT_GROUP_T;
  OriginQ O0 = make_origin(mk(0.f, 0.f, 0.f)), O1 = O0;
  DirQ D0, D1;
  D0.ndx = D0.ndy = D0.ndz = D0.od = 0.f; D1 = D0;
  bool live0 = false, live1 = false, exact0 = false, exact1 = false;
  if (s0 >= 0) {
    O0 = make_origin(slots[s0].qo);
    live0 = make_dir(D0, slots[s0].qo, slots[s0].rayD);
    exact0 = live0 && (p.noFilter || !(origin_filterable(O0) && dir_filterable(D0)));
  }
  if (s1 >= 0) {
    O1 = make_origin(slots[s1].qo);
    live1 = make_dir(D1, slots[s1].qo, slots[s1].rayD);
    exact1 = live1 && (p.noFilter || !(origin_filterable(O1) && dir_filterable(D1)));
  }
  constexpr unsigned m0 = mask_of_sub<2, G>(0), m1 = mask_of_sub<2, G>(1);
  const unsigned msk = pin(((live0 && !exact0) ? m0 : 0u) | ((live1 && !exact1) ? m1 : 0u));
  const Origin2 OO = pack_origin(O0, O1);
  const Dir2 DD = pack_dir(D0, D1);
  int cnt = 0;
  bool overflow = false;
  RT_TICK(1);
  if (!p.noFilter) {
    const uint32_t nPad = pin(w.nPad);
    for (uint32_t base = 0; base < nPad; base += G) {
      unsigned k = 0;
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const float4_ s = load_filt<USE_CONST>(w, cr, base + j);
        const f32x2 cx = pk1(s.x), cy = pk1(s.y), cz = pk1(s.z);
        const f32x2 b = bq2(DD, cx, cy, cz);
        f32x2 ch = fma2(OO.px, cx, pk1(s.w));
        ch = fma2(OO.py, cy, ch);
        ch = fma2(OO.pz, cz, ch);
        const f32x2 d = fma2(b, b, sub2(OO.nq, ch));     /* both rays: sign set <=> certain miss */
        k = __funnelshift_l(__float_as_uint(lo_of(d)), k, 1);
        k = __funnelshift_l(__float_as_uint(hi_of(d)), k, 1);
      }
      const unsigned comb = ~k & msk;
      if (comb) gather<2, G>(w, comb, base, cnt, overflow);
    }
  }
  float t0 = 1000.f, t1 = 1000.f;
  int h0 = -1, h1 = -1;
  const int maxc = __reduce_max_sync(RT_FULL, cnt);
  RT_TICK(2);
#pragma unroll 1
  for (int k = 0; k < maxc; ++k) {
    if (k < cnt && !overflow) {
      const uint32_t e = w.list[k * RT_BLOCK + w.tid];
      const uint32_t i = e & 0x3FFFu, sub = e >> 14;
      if (i < p.sc.n) {
        ctr.exactTests++;
        const Slot& q = slots[sub ? s1 : s0];
        const float t = ray_sphere_t(p.sc.geo[i], q.qo, q.rayD);
        if (t > 0.f) {       /* raytracer.h:166-188; strict <: first index wins ties */
          if (sub) { if (t < t1) { t1 = t; h1 = (int)i; } }
          else     { if (t < t0) { t0 = t; h0 = (int)i; } }
        }
      }
    }
  }
  if (overflow) { exact0 = live0; exact1 = live1; }
  if (s0 >= 0) { slots[s0].minT = t0; slots[s0].hitIdx = h0; }
  if (s1 >= 0) { slots[s1].minT = t1; slots[s1].hitIdx = h1; }
  if (exact0) ctr.exactTests += exact_all(p.sc, &slots[s0], 1u);
  if (exact1) ctr.exactTests += exact_all(p.sc, &slots[s1], 1u);
}

/* ---- shadow pass: the (up to four) shadow rays of one hit share their origin ------
 * A lane with no shadow batch waiting may bring a TRACE slot instead: a trace ray is the
 * same query with one direction, so it rides along for free and keeps the lane busy. */
template <bool USE_CONST, int ND>
__device__ __forceinline__ void pass_shadow(const TraceParams& p, const ConstRecords& cr, WarpCtx& w, Slot* slots, int s0,
                                            Counters& ctr) {
  constexpr int G = (ND == 4) ? RT_GROUP_S4 : RT_GROUP_S2;
  OriginQ O = make_origin(mk(0.f, 0.f, 0.f));
  DirQ D[ND];
  unsigned live = 0u, exact = 0u;
  bool asTrace = false;
#pragma unroll
  for (int k = 0; k < ND; ++k) { D[k].ndx = D[k].ndy = D[k].ndz = D[k].od = 0.f; }
  if (s0 >= 0) {
    /* the batch's rays (or the rider's single ray) go to the lane's shared-memory scratch:
     * resolve and advance read them back, the slot record does not carry them */
    asTrace = slots[s0].kind == K_TRACE;
    ShadowGeo g;
    V3 org;
    if (asTrace) {
      org = slots[s0].qo;
      g.d[0] = slots[s0].rayD; g.gap[0] = 0.f;
#pragma unroll
      for (int k = 1; k < RT_SHADOW_BATCH; ++k) { g.d[k] = mk(0.f, 0.f, 0.f); g.gap[k] = 0.f; }
    } else {
      org = slots[s0].P;
      shadow_geo(slots[s0], p.sc, g);
    }
#pragma unroll
    for (int k = 0; k < RT_SHADOW_BATCH; ++k) {
      w.geo[(4 * k + 0) * RT_BLOCK + w.tid] = g.d[k].x;
      w.geo[(4 * k + 1) * RT_BLOCK + w.tid] = g.d[k].y;
      w.geo[(4 * k + 2) * RT_BLOCK + w.tid] = g.d[k].z;
      w.geo[(4 * k + 3) * RT_BLOCK + w.tid] = g.gap[k];
    }
    O = make_origin(org);
    const bool ofil = origin_filterable(O);
    const int nd = slots[s0].ndirs;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      if (k < nd && make_dir(D[k], org, g.d[k])) {
        live |= 1u << k;
        if (p.noFilter || !(ofil && dir_filterable(D[k]))) exact |= 1u << k;
      }
    }
  }
  unsigned msk = 0u;
#pragma unroll
  for (int k = 0; k < ND; ++k) msk |= (((live & ~exact) >> k) & 1u) ? mask_of_sub<ND, G>(k) : 0u;
  msk = pin(msk);
  static_assert(ND % 2 == 0, "rays are processed as packed pairs");
  Dir2 DP[ND / 2];
#pragma unroll
  for (int k = 0; k < ND / 2; ++k) DP[k] = pack_dir(D[2 * k], D[2 * k + 1]);
  int cnt = 0;
  bool overflow = false;
  RT_TICK(1);
  if (!p.noFilter) {
    const uint32_t nPad = pin(w.nPad);
    for (uint32_t base = 0; base < nPad; base += G) {
      unsigned sk = 0u;
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const float4_ s = load_filt<USE_CONST>(w, cr, base + j);
        const f32x2 cx = pk1(s.x), cy = pk1(s.y), cz = pk1(s.z);
        const f32x2 e = pk1(ex_sub(O.nq, filter_ch(O, s)));   /* -q - ch: once per sphere, all rays share the origin */
#pragma unroll
        for (int k = 0; k < ND / 2; ++k) {
          const f32x2 b = bq2(DP[k], cx, cy, cz);
          const f32x2 d = fma2(b, b, e);
          sk = __funnelshift_l(__float_as_uint(lo_of(d)), sk, 1);
          sk = __funnelshift_l(__float_as_uint(hi_of(d)), sk, 1);
        }
      }
      const unsigned comb = ~sk & msk;
      if (comb) gather<ND, G>(w, comb, base, cnt, overflow);
    }
  }
  unsigned blocked = 0u;
  float t0 = 1000.f;
  int h0 = -1;
  const int maxc = __reduce_max_sync(RT_FULL, cnt);
  RT_TICK(2);
#pragma unroll 1
  for (int k = 0; k < maxc; ++k) {
    if (k < cnt && !overflow) {
      const uint32_t e = w.list[k * RT_BLOCK + w.tid];
      const uint32_t i = e & 0x3FFFu, sub = e >> 14;
      if (i < p.sc.n && !((blocked >> sub) & 1u)) {
        ctr.exactTests++;
        const V3 d = mk(w.geo[(4 * sub + 0) * RT_BLOCK + w.tid], w.geo[(4 * sub + 1) * RT_BLOCK + w.tid],
                        w.geo[(4 * sub + 2) * RT_BLOCK + w.tid]);
        const V3 org = asTrace ? slots[s0].qo : slots[s0].P;
        const float t = ray_sphere_t(p.sc.geo[i], org, d);
        if (t > 0.f) {
          if (asTrace) {          /* closest hit, raytracer.h:166-188 */
            if (t < t0) { t0 = t; h0 = (int)i; }
          } else if (t < 1000.f) { /* occluder iff |t d|^2 < gap, raytracer.h:291-304 (see resolve_shadow) */
            const V3 dist = vscale(t, d);
            if (vdot(dist, dist) < w.geo[(4 * sub + 3) * RT_BLOCK + w.tid]) blocked |= 1u << sub;
          }
        }
      }
    }
  }
  if (overflow) exact = live;
  if (s0 >= 0) {
    if (asTrace) { slots[s0].minT = t0; slots[s0].hitIdx = h0; }
    else slots[s0].blocked = blocked;
  }
  if (exact) ctr.exactTests += exact_all(p.sc, &slots[s0], exact);
}

/* ---- containment pass: up to two probe points per lane ---------------------------- */
template <bool USE_CONST>
__device__ __forceinline__ void pass_contain(const TraceParams& p, const ConstRecords& cr, WarpCtx& w, Slot* slots, int s0,
                                             int s1, Counters& ctr) {
  constexpr int G = RT_GROUP_C;
  OriginQ O0 = make_origin(mk(0.f, 0.f, 0.f)), O1 = O0;
  bool exact0 = false, exact1 = false;
  if (s0 >= 0) { O0 = make_origin(slots[s0].qo); exact0 = p.noFilter || !origin_filterable(O0); }
  if (s1 >= 0) { O1 = make_origin(slots[s1].qo); exact1 = p.noFilter || !origin_filterable(O1); }
  constexpr unsigned m0 = mask_of_sub<2, G>(0), m1 = mask_of_sub<2, G>(1);
  const unsigned msk = pin(((s0 >= 0 && !exact0) ? m0 : 0u) | ((s1 >= 0 && !exact1) ? m1 : 0u));
  const Origin2 OO = pack_origin(O0, O1);
  int cnt = 0;
  bool overflow = false;
  RT_TICK(1);
  if (!p.noFilter) {
    const uint32_t nPad = pin(w.nPad);
    for (uint32_t base = 0; base < nPad; base += G) {
      unsigned k = 0;
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const float4_ s = load_filt<USE_CONST>(w, cr, base + j);
        f32x2 ch = fma2(OO.px, pk1(s.x), pk1(s.w));
        ch = fma2(OO.py, pk1(s.y), ch);
        ch = fma2(OO.pz, pk1(s.z), ch);
        const f32x2 d = sub2(OO.nq, ch);                  /* both probes: sign set <=> certainly outside */
        k = __funnelshift_l(__float_as_uint(lo_of(d)), k, 1);
        k = __funnelshift_l(__float_as_uint(hi_of(d)), k, 1);
      }
      const unsigned comb = ~k & msk;
      if (comb) gather<2, G>(w, comb, base, cnt, overflow);
    }
  }
  int h0 = -1, h1 = -1;
  const int maxc = __reduce_max_sync(RT_FULL, cnt);
  RT_TICK(2);
#pragma unroll 1
  for (int k = 0; k < maxc; ++k) {
    if (k < cnt && !overflow) {
      const uint32_t e = w.list[k * RT_BLOCK + w.tid];
      const uint32_t i = e & 0x3FFFu, sub = e >> 14;
      if (i < p.sc.n) {
        ctr.exactTests++;
        const bool in = contains_exact(p.sc.geo[i], slots[sub ? s1 : s0].qo);
        if (in) {            /* raytracer.h:264: the first container in index order wins */
          if (sub) { if (h1 < 0) h1 = (int)i; }
          else     { if (h0 < 0) h0 = (int)i; }
        }
      }
    }
  }
  if (overflow) { exact0 = s0 >= 0; exact1 = s1 >= 0; }
  if (s0 >= 0) slots[s0].hitIdx = h0;
  if (s1 >= 0) slots[s1].hitIdx = h1;
  if (exact0) ctr.exactTests += exact_all(p.sc, &slots[s0], 1u);
  if (exact1) ctr.exactTests += exact_all(p.sc, &slots[s1], 1u);
}

/* ======================================================================================
 * Accelerated mode (optional; "Cluster filter" in rt_core.cuh).  Each pass runs the SAME packed
 * loop over the clusters' bounding records (an eighth of the sphere count), collects the
 * (sub-query, cluster) pairs it cannot rule out, tests those clusters' members with the
 * ordinary per-sphere filter and resolves the surviving spheres exactly.  The lists are
 * flushed whenever a lane's list could overflow, so no input makes this mode fall back to
 * the all-spheres resolve except what the brute-force mode also sends there.
 * ====================================================================================== */
template <int ND, int G>
__device__ __forceinline__ void gather1(const WarpCtx& w, unsigned comb, uint32_t base, int& cnt) {
  while (comb) {                      /* the caller made sure the list has room */
    const int b = 31 - __clz(comb);
    comb &= ~(1u << b);
    const uint32_t idx = (uint32_t)(G * ND - 1 - b);
    w.list1[cnt * RT_BLOCK + w.tid] = (unsigned short)(((idx % ND) << 14) | (base + idx / ND));
    ++cnt;
  }
}

/* members of cluster `cl` the per-sphere filter cannot rule out for one query (bit 7-m = member m) */
template <bool HAS_DIR>
__device__ __forceinline__ unsigned member_bits(const WarpCtx& w, uint32_t cl, const OriginQ& O, const DirQ& D) {
  unsigned bits = 0u;
#pragma unroll
  for (int m = 0; m < RT_CLUSTER; ++m) {
    const float4 v = w.mfilt[cl * RT_CLUSTER + m];
    float4_ s; s.x = v.x; s.y = v.y; s.z = v.z; s.w = v.w;
    const float ch = filter_ch(O, s);
    const float d = HAS_DIR ? filter_ray(O, D, ch, s) : filter_point(O, ch);
    bits = __funnelshift_l(__float_as_uint(d), bits, 1);
  }
  return ~bits & ((1u << RT_CLUSTER) - 1u);
}
/* append the flagged members of one cluster to the sphere list */
__device__ __forceinline__ void gather2(const WarpCtx& w, unsigned bits, uint32_t cl, uint32_t sub, int& cnt2,
                                        bool& overflow) {
  while (bits) {
    const int b = 31 - __clz(bits);
    bits &= ~(1u << b);
    if (cnt2 < RT_LIST_MAX) {
      w.list[cnt2 * RT_BLOCK + w.tid] =
          (unsigned short)((sub << 14) | w.midx[cl * RT_CLUSTER + (uint32_t)(RT_CLUSTER - 1 - b)]);
      ++cnt2;
    } else {
      overflow = true;
    }
  }
}

template <bool USE_CONST>
__device__ __forceinline__ void pass_trace_accel(const TraceParams& p, const ConstRecords& cr, WarpCtx& w, Slot* slots, int s0,
                                                 int s1, Counters& ctr) {
  constexpr int G = RT_GROUP_TA;
  OriginQ O0 = make_origin(mk(0.f, 0.f, 0.f)), O1 = O0;
  DirQ D0, D1;
  D0.ndx = D0.ndy = D0.ndz = D0.od = 0.f; D1 = D0;
  bool live0 = false, live1 = false, exact0 = false, exact1 = false;
  if (s0 >= 0) {
    O0 = make_origin(slots[s0].qo);
    live0 = make_dir(D0, slots[s0].qo, slots[s0].rayD);
    exact0 = live0 && (p.noFilter || !(origin_filterable(O0) && dir_filterable(D0)));
  }
  if (s1 >= 0) {
    O1 = make_origin(slots[s1].qo);
    live1 = make_dir(D1, slots[s1].qo, slots[s1].rayD);
    exact1 = live1 && (p.noFilter || !(origin_filterable(O1) && dir_filterable(D1)));
  }
  constexpr unsigned m0 = mask_of_sub<2, G>(0), m1 = mask_of_sub<2, G>(1);
  const unsigned msk = pin(((live0 && !exact0) ? m0 : 0u) | ((live1 && !exact1) ? m1 : 0u));
  const Origin2 OO = pack_origin(cluster_origin(O0), cluster_origin(O1));
  const Dir2 DD = pack_dir(D0, D1);
  float t0 = 1000.f, t1 = 1000.f;
  int h0 = -1, h1 = -1;
  bool overflow = false;
  RT_TICK(1);
  const uint32_t nPad = pin(p.noFilter ? 0u : w.nPad);
  uint32_t base = 0;
  for (;;) {
    int cnt = 0;
    bool full = false;
    for (; base < nPad; base += G) {
      unsigned k = 0;
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const float4_ s = load_filt<USE_CONST>(w, cr, base + j);
        const f32x2 cx = pk1(s.x), cy = pk1(s.y), cz = pk1(s.z);
        const f32x2 b = bq2(DD, cx, cy, cz);
        f32x2 ch = fma2(OO.px, cx, pk1(s.w));
        ch = fma2(OO.py, cy, ch);
        ch = fma2(OO.pz, cz, ch);
        const f32x2 d = fma2(b, b, sub2(OO.nq, ch));
        k = __funnelshift_l(__float_as_uint(lo_of(d)), k, 1);
        k = __funnelshift_l(__float_as_uint(hi_of(d)), k, 1);
      }
      const unsigned comb = ~k & msk;
      if (__any_sync(RT_FULL, cnt + __popc(comb) > (int)p.list1Max)) { full = true; break; }
      if (comb) gather1<2, G>(w, comb, base, cnt);
    }
    RT_TICK(2);
    /* members of the surviving clusters */
    /* members of the surviving clusters */
    int cnt2 = 0;
    const int maxc = __reduce_max_sync(RT_FULL, cnt);
#pragma unroll 1
    for (int k = 0; k < maxc; ++k) {
      if (k < cnt) {
        const uint32_t e = w.list1[k * RT_BLOCK + w.tid];
        const uint32_t cl = e & 0x3FFFu, sub = e >> 14;
        OriginQ O = O0; DirQ D = D0;
        if (sub) { O = O1; D = D1; }
        gather2(w, member_bits<true>(w, cl, O, D), cl, sub, cnt2, overflow);
      }
    }
    const int maxc2 = __reduce_max_sync(RT_FULL, cnt2);
#pragma unroll 1
    for (int k = 0; k < maxc2; ++k) {
      if (k < cnt2 && !overflow) {
        const uint32_t e = w.list[k * RT_BLOCK + w.tid];
        const uint32_t i = e & 0x3FFFu, sub = e >> 14;
        if (i < p.sc.n) {
          ctr.exactTests++;
          const Slot& q = slots[sub ? s1 : s0];
          const float t = ray_sphere_t(p.sc.geo[i], q.qo, q.rayD);
          if (t > 0.f) {       /* raytracer.h:166-188: closest hit, the first index wins ties */
            if (sub) { if (t < t1 || (t == t1 && (int)i < h1)) { t1 = t; h1 = (int)i; } }
            else     { if (t < t0 || (t == t0 && (int)i < h0)) { t0 = t; h0 = (int)i; } }
          }
        }
      }
    }
    RT_TICK(3);
    if (!full) break;
  }
  if (overflow) { exact0 = live0; exact1 = live1; }
  if (s0 >= 0) { slots[s0].minT = t0; slots[s0].hitIdx = h0; }
  if (s1 >= 0) { slots[s1].minT = t1; slots[s1].hitIdx = h1; }
  if (exact0) ctr.exactTests += exact_all(p.sc, &slots[s0], 1u);
  if (exact1) ctr.exactTests += exact_all(p.sc, &slots[s1], 1u);
}

template <bool USE_CONST, int ND>
__device__ __forceinline__ void pass_shadow_accel(const TraceParams& p, const ConstRecords& cr, WarpCtx& w, Slot* slots, int s0,
                                                  Counters& ctr) {
  constexpr int G = (ND == 4) ? RT_GROUP_S4A : RT_GROUP_S2;
  OriginQ O = make_origin(mk(0.f, 0.f, 0.f));
  DirQ D[ND];
  unsigned live = 0u, exact = 0u;
  bool asTrace = false;
#pragma unroll
  for (int k = 0; k < ND; ++k) { D[k].ndx = D[k].ndy = D[k].ndz = D[k].od = 0.f; }
  if (s0 >= 0) {
    asTrace = slots[s0].kind == K_TRACE;
    ShadowGeo g;
    V3 org;
    if (asTrace) {
      org = slots[s0].qo;
      g.d[0] = slots[s0].rayD; g.gap[0] = 0.f;
#pragma unroll
      for (int k = 1; k < RT_SHADOW_BATCH; ++k) { g.d[k] = mk(0.f, 0.f, 0.f); g.gap[k] = 0.f; }
    } else {
      org = slots[s0].P;
      shadow_geo(slots[s0], p.sc, g);
    }
#pragma unroll
    for (int k = 0; k < RT_SHADOW_BATCH; ++k) {
      w.geo[(4 * k + 0) * RT_BLOCK + w.tid] = g.d[k].x;
      w.geo[(4 * k + 1) * RT_BLOCK + w.tid] = g.d[k].y;
      w.geo[(4 * k + 2) * RT_BLOCK + w.tid] = g.d[k].z;
      w.geo[(4 * k + 3) * RT_BLOCK + w.tid] = g.gap[k];
    }
    O = make_origin(org);
    const bool ofil = origin_filterable(O);
    const int nd = slots[s0].ndirs;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      if (k < nd && make_dir(D[k], org, g.d[k])) {
        live |= 1u << k;
        if (p.noFilter || !(ofil && dir_filterable(D[k]))) exact |= 1u << k;
      }
    }
  }
  unsigned msk = 0u;
#pragma unroll
  for (int k = 0; k < ND; ++k) msk |= (((live & ~exact) >> k) & 1u) ? mask_of_sub<ND, G>(k) : 0u;
  msk = pin(msk);
  Dir2 DP[ND / 2];
#pragma unroll
  for (int k = 0; k < ND / 2; ++k) DP[k] = pack_dir(D[2 * k], D[2 * k + 1]);
  const OriginQ OC = cluster_origin(O);
  unsigned blocked = 0u;
  float t0 = 1000.f;
  int h0 = -1;
  bool overflow = false;
  RT_TICK(1);
  const uint32_t nPad = pin(p.noFilter ? 0u : w.nPad);
  uint32_t base = 0;
  for (;;) {
    int cnt = 0;
    bool full = false;
    for (; base < nPad; base += G) {
      unsigned sk = 0u;
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const float4_ s = load_filt<USE_CONST>(w, cr, base + j);
        const f32x2 cx = pk1(s.x), cy = pk1(s.y), cz = pk1(s.z);
        const f32x2 e = pk1(ex_sub(OC.nq, filter_ch(OC, s)));
#pragma unroll
        for (int k = 0; k < ND / 2; ++k) {
          const f32x2 b = bq2(DP[k], cx, cy, cz);
          const f32x2 d = fma2(b, b, e);
          sk = __funnelshift_l(__float_as_uint(lo_of(d)), sk, 1);
          sk = __funnelshift_l(__float_as_uint(hi_of(d)), sk, 1);
        }
      }
      const unsigned comb = ~sk & msk;
      if (__any_sync(RT_FULL, cnt + __popc(comb) > (int)p.list1Max)) { full = true; break; }
      if (comb) gather1<ND, G>(w, comb, base, cnt);
    }
    RT_TICK(2);
    int cnt2 = 0;
    const int maxc = __reduce_max_sync(RT_FULL, cnt);
#pragma unroll 1
    for (int k = 0; k < maxc; ++k) {
      if (k < cnt) {
        const uint32_t e = w.list1[k * RT_BLOCK + w.tid];
        const uint32_t cl = e & 0x3FFFu, sub = e >> 14;
        DirQ Ds = D[0];
#pragma unroll
        for (int j = 1; j < ND; ++j) if (sub == (uint32_t)j) Ds = D[j];
        gather2(w, member_bits<true>(w, cl, O, Ds), cl, sub, cnt2, overflow);
      }
    }
    const int maxc2 = __reduce_max_sync(RT_FULL, cnt2);
#pragma unroll 1
    for (int k = 0; k < maxc2; ++k) {
      if (k < cnt2 && !overflow) {
        const uint32_t e = w.list[k * RT_BLOCK + w.tid];
        const uint32_t i = e & 0x3FFFu, sub = e >> 14;
        if (i < p.sc.n && !((blocked >> sub) & 1u)) {
          ctr.exactTests++;
          const V3 d = mk(w.geo[(4 * sub + 0) * RT_BLOCK + w.tid], w.geo[(4 * sub + 1) * RT_BLOCK + w.tid],
                          w.geo[(4 * sub + 2) * RT_BLOCK + w.tid]);
          const V3 org = asTrace ? slots[s0].qo : slots[s0].P;
          const float t = ray_sphere_t(p.sc.geo[i], org, d);
          if (t > 0.f) {
            if (asTrace) {
              if (t < t0 || (t == t0 && (int)i < h0)) { t0 = t; h0 = (int)i; }
            } else if (t < 1000.f) {
              const V3 dist = vscale(t, d);
              if (vdot(dist, dist) < w.geo[(4 * sub + 3) * RT_BLOCK + w.tid]) blocked |= 1u << sub;
            }
          }
        }
      }
    }
    RT_TICK(3);
    if (!full) break;
  }
  if (overflow) exact = live;
  if (s0 >= 0) {
    if (asTrace) { slots[s0].minT = t0; slots[s0].hitIdx = h0; }
    else slots[s0].blocked = blocked;
  }
  if (exact) ctr.exactTests += exact_all(p.sc, &slots[s0], exact);
}

template <bool USE_CONST>
__device__ __forceinline__ void pass_contain_accel(const TraceParams& p, const ConstRecords& cr, WarpCtx& w, Slot* slots, int s0,
                                                   int s1, Counters& ctr) {
  constexpr int G = RT_GROUP_CA;
  OriginQ O0 = make_origin(mk(0.f, 0.f, 0.f)), O1 = O0;
  bool exact0 = false, exact1 = false;
  if (s0 >= 0) { O0 = make_origin(slots[s0].qo); exact0 = p.noFilter || !origin_filterable(O0); }
  if (s1 >= 0) { O1 = make_origin(slots[s1].qo); exact1 = p.noFilter || !origin_filterable(O1); }
  constexpr unsigned m0 = mask_of_sub<2, G>(0), m1 = mask_of_sub<2, G>(1);
  const unsigned msk = pin(((s0 >= 0 && !exact0) ? m0 : 0u) | ((s1 >= 0 && !exact1) ? m1 : 0u));
  const Origin2 OO = pack_origin(cluster_origin(O0), cluster_origin(O1));
  int h0 = -1, h1 = -1;
  bool overflow = false;
  RT_TICK(1);
  const uint32_t nPad = pin(p.noFilter ? 0u : w.nPad);
  uint32_t base = 0;
  DirQ none;
  none.ndx = none.ndy = none.ndz = none.od = 0.f;
  for (;;) {
    int cnt = 0;
    bool full = false;
    for (; base < nPad; base += G) {
      unsigned k = 0;
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const float4_ s = load_filt<USE_CONST>(w, cr, base + j);
        f32x2 ch = fma2(OO.px, pk1(s.x), pk1(s.w));
        ch = fma2(OO.py, pk1(s.y), ch);
        ch = fma2(OO.pz, pk1(s.z), ch);
        const f32x2 d = sub2(OO.nq, ch);
        k = __funnelshift_l(__float_as_uint(lo_of(d)), k, 1);
        k = __funnelshift_l(__float_as_uint(hi_of(d)), k, 1);
      }
      const unsigned comb = ~k & msk;
      if (__any_sync(RT_FULL, cnt + __popc(comb) > (int)p.list1Max)) { full = true; break; }
      if (comb) gather1<2, G>(w, comb, base, cnt);
    }
    RT_TICK(2);
    int cnt2 = 0;
    const int maxc = __reduce_max_sync(RT_FULL, cnt);
#pragma unroll 1
    for (int k = 0; k < maxc; ++k) {
      if (k < cnt) {
        const uint32_t e = w.list1[k * RT_BLOCK + w.tid];
        const uint32_t cl = e & 0x3FFFu, sub = e >> 14;
        OriginQ O = O0;
        if (sub) O = O1;
        gather2(w, member_bits<false>(w, cl, O, none), cl, sub, cnt2, overflow);
      }
    }
    const int maxc2 = __reduce_max_sync(RT_FULL, cnt2);
#pragma unroll 1
    for (int k = 0; k < maxc2; ++k) {
      if (k < cnt2 && !overflow) {
        const uint32_t e = w.list[k * RT_BLOCK + w.tid];
        const uint32_t i = e & 0x3FFFu, sub = e >> 14;
        if (i < p.sc.n) {
          ctr.exactTests++;
          const bool in = contains_exact(p.sc.geo[i], slots[sub ? s1 : s0].qo);
          if (in) {            /* raytracer.h:264: the first container in index order wins */
            if (sub) { if (h1 < 0 || (int)i < h1) h1 = (int)i; }
            else     { if (h0 < 0 || (int)i < h0) h0 = (int)i; }
          }
        }
      }
    }
    RT_TICK(3);
    if (!full) break;
  }
  if (overflow) { exact0 = s0 >= 0; exact1 = s1 >= 0; }
  if (s0 >= 0) slots[s0].hitIdx = h0;
  if (s1 >= 0) slots[s1].hitIdx = h1;
  if (exact0) ctr.exactTests += exact_all(p.sc, &slots[s0], 1u);
  if (exact1) ctr.exactTests += exact_all(p.sc, &slots[s1], 1u);
}

/* Pull the state words of a slot that is about to be served into L1 while the sphere loop
 * runs: the per-lane slots live in local memory (one 128-byte line per word per warp), far
 * more than L1 holds across all resident warps, so without this the O(1) code after the loop
 * waits on L2 for every field it touches. */
__device__ __forceinline__ void prefetch_slot(const Slot& s) {
  const float* w = reinterpret_cast<const float*>(&s);
  constexpr int WORDS = (int)((sizeof(Slot) - sizeof(Frame) * RT_MAX_STACK) / sizeof(float));
#pragma unroll 1
  for (int i = 0; i < WORDS; ++i)
    asm volatile("{ .reg .u64 la; cvta.to.local.u64 la, %0; prefetch.local.L1 [la]; }" ::"l"(w + i));
}

/* Advance one served slot; store the pixel when it completes.  The slot record is copied
 * into registers first (one batch of independent local loads instead of a dependent
 * load/store chain through L2) and written back once.  Returns the slot's new
 * (kind | ndirs << 4) tag for the lane's register-resident census. */
__device__ __forceinline__ uint32_t advance_slot(const TraceParams& p, Slot* slot, Frame* stack, Counters& ctr,
                                                 float& laneMax, const ShadowGeo* sg) {
  Slot s = *slot;
  if (advance(s, stack, ctr, p.sc, p.cam, sg)) {
    const V3 v = sample_value(s, p.cam);
    if (p.spp == 1u) {
      /* one sample per pixel: the pixel is 0 + sample (main.cpp:420,446) */
      const V3 a = vadd(mk(0.f, 0.f, 0.f), v);
      p.fb[s.pixel] = make_float4(a.x, a.y, a.z, 1.f);
      if (a.x > laneMax) laneMax = a.x;   /* algebra.h:74-82, NaN skipped */
      if (a.y > laneMax) laneMax = a.y;
      if (a.z > laneMax) laneMax = a.z;
    } else {
      p.samples[s.pixel] = make_float4(v.x, v.y, v.z, 1.f);
    }
    s.pixel = RT_NO_PIXEL;
    s.kind = K_NULL;
    s.ndirs = 0;
  }
  *slot = s;
  return (uint32_t)s.kind | ((uint32_t)s.ndirs << 4);
}

template <bool USE_CONST, int NSLOTS, bool ACCEL>
__device__ __forceinline__ void trace_body(const TraceParams& p, const ConstRecords& cr) {
  static_assert(!(ACCEL && USE_CONST), "the accelerated mode stages its records in shared memory");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  /* layout: [mbarrier 16 B][filter records nPad*16 B (shared staging only)][lists][shadow-ray scratch];
   * accelerated mode: [mbarrier][cluster records | member records | member indices][list1][list][scratch] */
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float4* sFilt = reinterpret_cast<float4*>(smem_raw + 16);
  const uint32_t filtBytes = USE_CONST ? 0u : ACCEL ? p.sc.ncPad * (16u + RT_CLUSTER * 18u) : p.sc.nPad * 16u;

  const uint32_t tid = threadIdx.x, lane = tid & 31u;

  if (!USE_CONST) {
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && filtBytes) {
      mbar_expect_tx(bar, filtBytes);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(ACCEL ? p.sc.cfilt : p.sc.filt);
      unsigned char* dstp = reinterpret_cast<unsigned char*>(sFilt);
      for (uint32_t off = 0; off < filtBytes; off += 32768u) {
        const uint32_t n = (filtBytes - off < 32768u) ? (filtBytes - off) : 32768u;
        tma_bulk_g2s(dstp + off, src + off, n, bar);
      }
    }
    if (filtBytes) mbar_wait(bar, 0);
  }

  WarpCtx w;
  w.filt = sFilt;
  w.mfilt = sFilt + p.sc.ncPad;
  w.midx = reinterpret_cast<const unsigned short*>(sFilt + (size_t)p.sc.ncPad * (1u + RT_CLUSTER));
  const uint32_t list1Bytes = ACCEL ? p.list1Max * RT_BLOCK * (uint32_t)sizeof(unsigned short) : 0u;
  w.list1 = reinterpret_cast<unsigned short*>(smem_raw + 16 + filtBytes);
  w.list = reinterpret_cast<unsigned short*>(smem_raw + 16 + filtBytes + list1Bytes);
  w.geo = reinterpret_cast<float*>(smem_raw + 16 + filtBytes + list1Bytes + RT_LIST_MAX * RT_BLOCK * sizeof(unsigned short));
  w.tid = tid;
  w.nPad = ACCEL ? p.sc.ncPad : p.sc.nPad;
#ifdef RT_PHASE_TIMING
  w.t0 = w.tDry = clock64();
  for (int i = 0; i < 6; ++i) w.phase[i] = 0;
#endif

  Slot slots[NSLOTS];
  Frame stacks[NSLOTS * RT_MAX_STACK];
#pragma unroll 1
  for (int k = 0; k < NSLOTS; ++k) { slots[k].kind = K_NULL; slots[k].pixel = RT_NO_PIXEL; slots[k].ndirs = 0; }
  uint32_t tags = 0u;   /* census kept in a register: byte k = kind | ndirs << 4 of slot k */
  Counters ctr;
  ctr.rays = ctr.shadow = ctr.containQ = ctr.containT = ctr.exactTests = ctr.samples = ctr.nullRays = 0;
  float laneMax = 0.f;
  uint32_t passT = 0, passS2 = 0, passS4 = 0, passC = 0;   /* passes by kind (warp-uniform) */
  uint32_t servedT = 0, servedS = 0, servedC = 0;            /* sub-queries of this lane served */

  uint32_t wbase = 0, wend = 0;          /* warp-uniform slice of the tile queue */
  bool queueDry = false;

  for (;;) {
    /* ---- refill ---- */
#pragma unroll 1
    for (int k = 0; k < NSLOTS && !queueDry; ++k) {
      bool need = (((tags >> (8 * k)) & 0xFu) == (uint32_t)K_NULL);
      while (!queueDry) {
        const unsigned m = __ballot_sync(RT_FULL, need);
        if (m == 0) break;
        if (wbase >= wend) {
          uint32_t b = 0;
          if (lane == 0) b = atomicAdd(p.workCounter, p.chunk);
          b = __shfl_sync(RT_FULL, b, 0);
          if (b >= p.totalWork) {
            queueDry = true;
#ifdef RT_PHASE_TIMING
            w.tDry = clock64();
#endif
            break;
          }
          wbase = b;
          wend = (b + p.chunk < p.totalWork) ? b + p.chunk : p.totalWork;
        }
        const uint32_t avail = wend - wbase;
        const uint32_t rank = __popc(m & ((1u << lane) - 1u));
        if (need && rank < avail) {
          uint32_t gx, gy, dst;
          int si, sj;
          if (work_to_task(p, wbase + rank, gx, gy, dst, si, sj)) {
            start_task(slots[k], ctr, p.cam, gx, gy, dst, si, sj);
            need = false;
            tags = (tags & ~(0xFFu << (8 * k))) | (((uint32_t)K_TRACE | (1u << 4)) << (8 * k));
          }
        }
        const uint32_t cnt = __popc(m);
        wbase += (cnt < avail) ? cnt : avail;
      }
    }

    /* ---- vote ---- */
    int t0 = -1, t1 = -1, s0 = -1, c0 = -1, c1 = -1, nd = 0;
#pragma unroll
    for (int k = 0; k < NSLOTS; ++k) {
      const uint32_t tg = (tags >> (8 * k)) & 0xFFu;
      const int kind = (int)(tg & 0xFu);
      if (kind == K_TRACE) { if (t0 < 0) t0 = k; else if (t1 < 0) t1 = k; }
      else if (kind == K_SHADOW) { if (s0 < 0) { s0 = k; nd = (int)(tg >> 4); } }
      else if (kind == K_CONTAIN) { if (c0 < 0) c0 = k; else if (c1 < 0) c1 = k; }
    }
    /* a shadow pass also takes one trace ray from lanes that have no shadow batch waiting */
    const int sOrT = (s0 >= 0) ? s0 : t0;
    const int ndS = (s0 >= 0) ? nd : (t0 >= 0 ? 1 : 0);
    const int ndMax = __reduce_max_sync(RT_FULL, nd);
    const unsigned nT = __reduce_add_sync(RT_FULL, (unsigned)((t0 >= 0) + (t1 >= 0)));
    const unsigned nS = __reduce_add_sync(RT_FULL, (unsigned)ndS);
    const unsigned nC = __reduce_add_sync(RT_FULL, (unsigned)((c0 >= 0) + (c1 >= 0)));
    if ((nT | nS | nC) == 0u) {
      if (queueDry) break;
      continue;
    }
    /* serve the kind that fills the largest share of its pass: capacity 64 sub-queries for
     * trace / contain passes, 32 x (2 or 4) for a shadow pass */
    const unsigned capS = (ndMax <= 2) ? 64u : 128u;
    int sv0, sv1;
    const bool anyS = ndMax > 0;
    const int mode = (anyS && nS * 64u >= nT * capS && nS * 64u >= nC * capS) ? K_SHADOW
                     : (nT >= nC) ? K_TRACE : K_CONTAIN;
    sv0 = (mode == K_SHADOW) ? sOrT : (mode == K_TRACE) ? t0 : c0;
    sv1 = (mode == K_SHADOW) ? -1 : (mode == K_TRACE) ? t1 : c1;
    if (p.prefetch) {
      if (sv0 >= 0) prefetch_slot(slots[sv0]);
      if (sv1 >= 0) prefetch_slot(slots[sv1]);
    }
    if (mode == K_SHADOW) {
      if (ndMax <= 2 && !(ACCEL && RT_ACCEL_ONE_SHADOW)) {
        if (ACCEL) pass_shadow_accel<USE_CONST, 2>(p, cr, w, slots, sv0, ctr); else pass_shadow<USE_CONST, 2>(p, cr, w, slots, sv0, ctr);
        passS2++;
      } else {
        if (ACCEL) pass_shadow_accel<USE_CONST, 4>(p, cr, w, slots, sv0, ctr); else pass_shadow<USE_CONST, 4>(p, cr, w, slots, sv0, ctr);
        passS4++;
      }
      if (s0 >= 0) servedS += (unsigned)nd; else servedT += (unsigned)ndS;
    } else if (mode == K_TRACE) {
      if (ACCEL) pass_trace_accel<USE_CONST>(p, cr, w, slots, t0, t1, ctr); else pass_trace<USE_CONST>(p, cr, w, slots, t0, t1, ctr);
      passT++; servedT += (unsigned)((t0 >= 0) + (t1 >= 0));
    } else {
      if (ACCEL) pass_contain_accel<USE_CONST>(p, cr, w, slots, c0, c1, ctr); else pass_contain<USE_CONST>(p, cr, w, slots, c0, c1, ctr);
      passC++; servedC += (unsigned)((c0 >= 0) + (c1 >= 0));
    }
    RT_TICK(3);
    /* ---- advance the served slots (one code instance, same kind across the warp) ---- */
    ShadowGeo sg;
    if (mode == K_SHADOW && sv0 >= 0) {       /* the rays of the batch just served */
#pragma unroll
      for (int k = 0; k < RT_SHADOW_BATCH; ++k) {
        sg.d[k] = mk(w.geo[(4 * k + 0) * RT_BLOCK + tid], w.geo[(4 * k + 1) * RT_BLOCK + tid],
                     w.geo[(4 * k + 2) * RT_BLOCK + tid]);
        sg.gap[k] = w.geo[(4 * k + 3) * RT_BLOCK + tid];
      }
    }
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
      const int sv = r ? sv1 : sv0;
      if (sv >= 0) {
        const uint32_t tg = advance_slot(p, &slots[sv], &stacks[sv * RT_MAX_STACK], ctr, laneMax, &sg);
        tags = (tags & ~(0xFFu << (8 * sv))) | (tg << (8 * sv));
      }
    }
    RT_TICK(4);
  }

#ifdef RT_PHASE_TIMING
  if (lane == 0)
  {
    w.phase[0] = w.phase[5] = clock64() - w.tDry;
    for (int i = 0; i < 5; ++i) atomicAdd(&p.counters[16 + i], (unsigned long long)w.phase[i]);
    atomicMax(&p.counters[16 + 5], (unsigned long long)w.phase[5]);
  }
#endif
  /* ---- per-warp reductions ---- */
  unsigned mb = __float_as_uint(laneMax);   /* laneMax >= 0: uint order == float order */
  mb = __reduce_max_sync(RT_FULL, mb);
  unsigned long long v[10];
  v[0] = ctr.rays; v[1] = ctr.shadow; v[2] = ctr.containQ; v[3] = ctr.containT;
  v[4] = ctr.exactTests; v[5] = ctr.samples; v[6] = ctr.nullRays;
  v[7] = servedT; v[8] = servedS; v[9] = servedC;
#pragma unroll
  for (int i = 0; i < 10; ++i)
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(RT_FULL, v[i], o);
  if (lane == 0) {
    if (mb) atomicMax(p.maxBits, mb);
#pragma unroll
    for (int i = 0; i < 10; ++i) atomicAdd(&p.counters[i], v[i]);
    /* sub-query capacity offered by this warp's passes (32 lanes x 2 / ND / 2 per pass) */
    atomicAdd(&p.counters[10], 64ull * passT + 64ull * passS2 + 128ull * passS4 + 64ull * passC);
    atomicAdd(&p.counters[11], 1ull * passT + 1ull * passS2 + 1ull * passS4 + 1ull * passC);
    atomicAdd(&p.counters[12], (unsigned long long)passT);
    atomicAdd(&p.counters[13], (unsigned long long)passS2);
    atomicAdd(&p.counters[14], (unsigned long long)passS4);
    atomicAdd(&p.counters[15], (unsigned long long)passC);
  }
}

/* The two entry points: records staged in shared memory by TMA (every scene size, and the
 * accelerated mode), or read from the launch's constant bank (<= RT_CONST_MAX_SPHERES records). */
template <int MIN_BLOCKS, int NSLOTS, bool ACCEL>
__global__ void __launch_bounds__(RT_BLOCK, MIN_BLOCKS) trace_kernel(const __grid_constant__ TraceParams p) {
  trace_body<false, NSLOTS, ACCEL>(p, *reinterpret_cast<const ConstRecords*>(0));   /* never read without USE_CONST */
}
template <int MIN_BLOCKS, int NSLOTS>
__global__ void __launch_bounds__(RT_BLOCK, MIN_BLOCKS) trace_kernel_const(const __grid_constant__ TraceParams p,
                                                                          const __grid_constant__ ConstRecords c) {
  trace_body<true, NSLOTS, false>(p, c);
}

/* Sum each pixel's samples in the reference's order (main.cpp:430-447) and take the frame's
 * NaN-skipping maximum (algebra.h:68-91).  HBM-bound: 16 B x spp read + 16 B written per pixel. */
__global__ void combine_kernel(const float4* __restrict__ samples, float4* __restrict__ fb, uint32_t npix,
                               uint32_t spp, unsigned int* __restrict__ maxBits) {
  float m = 0.f;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t px = blockIdx.x * blockDim.x + threadIdx.x; px < npix; px += stride) {
    V3 acc = mk(0.f, 0.f, 0.f);
    for (uint32_t k = 0; k < spp; ++k) {
      const float4 v = samples[(size_t)px * spp + k];
      acc = vadd(acc, mk(v.x, v.y, v.z));
    }
    fb[px] = make_float4(acc.x, acc.y, acc.z, 1.f);
    if (acc.x > m) m = acc.x;
    if (acc.y > m) m = acc.y;
    if (acc.z > m) m = acc.z;
  }
  unsigned mb = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(m));
  if ((threadIdx.x & 31u) == 0 && mb) atomicMax(maxBits, mb);
}

/* float4 framebuffer -> packed 12-byte pixels (the reference's `Vec dst[]`, .cl:972) */
__global__ void pack_kernel(const float4* __restrict__ fb, float* __restrict__ out, uint32_t npix) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nflt = (size_t)npix * 3u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nflt; i += stride) {
    const size_t px = i / 3u, c = i - px * 3u;
    const float* f = reinterpret_cast<const float*>(fb + px);
    out[i] = f[c];
  }
}

/* main.cpp:71-76 with g++/x86-64 cast semantics (cvttss2si, low byte) */
__device__ __forceinline__ unsigned quantise1(float v, float maxv) {
  const float clipped = (v < 1.f) ? v : 1.f;
  const float s = ex_div(ex_mul(clipped, 255.f), maxv);
  int iv;
  if (s >= -2147483648.0f && s < 2147483648.0f) iv = __float2int_rz(s); else iv = (int)0x80000000;
  return (unsigned)iv & 0xFFu;
}

/* 4 pixels (12 bytes, three aligned 32-bit stores) per thread */
__global__ void quantise_kernel(const float4* __restrict__ fb, uint32_t* __restrict__ out,
                                uint32_t npix, const unsigned int* __restrict__ maxBits,
                                float maxOverride) {
  float maxv = (maxOverride > 0.f) ? maxOverride : __uint_as_float(*maxBits);
  if (maxv == 0.f) maxv = 1.f;   /* algebra.h:86-88 */
  const uint32_t quads = (npix + 3u) >> 2;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t qd = blockIdx.x * blockDim.x + threadIdx.x; qd < quads; qd += stride) {
    unsigned b[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t px = qd * 4u + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (px < npix) v = fb[px];
      b[3 * k + 0] = quantise1(v.x, maxv);
      b[3 * k + 1] = quantise1(v.y, maxv);
      b[3 * k + 2] = quantise1(v.z, maxv);
    }
    const uint32_t w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    const uint32_t w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    const uint32_t w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
    if (qd * 4u + 3u < npix) {
      out[qd * 3u + 0] = w0; out[qd * 3u + 1] = w1; out[qd * 3u + 2] = w2;
    } else {
      unsigned char* o8 = reinterpret_cast<unsigned char*>(out) + (size_t)qd * 12u;
      const uint32_t rem = npix - qd * 4u;
      for (uint32_t k = 0; k < rem * 3u; ++k) o8[k] = (unsigned char)b[k];
    }
  }
}

/* Multi-GPU assembly: `gathered` holds nShards blocks of shardPitch bytes, block g =
 * the RGB8 rows of shard g packed in increasing row order (rt_cuda_render_strips).
 * One 16-byte vector per thread where alignment allows. */
__global__ void assemble_rgb8_kernel(const unsigned char* __restrict__ gathered,
                                     unsigned char* __restrict__ out, uint32_t width, uint32_t height,
                                     uint32_t stripRows, uint32_t nShards, size_t shardPitch) {
  const uint32_t rowBytes = width * 3u;
  for (uint32_t row = blockIdx.x; row < height; row += gridDim.x) {
    const uint32_t strip = row / stripRows;
    const uint32_t shard = strip % nShards;
    const uint32_t localRow = (strip / nShards) * stripRows + (row - strip * stripRows);
    const unsigned char* src = gathered + (size_t)shard * shardPitch + (size_t)localRow * rowBytes;
    unsigned char* dst = out + (size_t)row * rowBytes;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0) {
      const uint32_t nvec = rowBytes >> 4;
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) d4[i] = s4[i];
      for (uint32_t i = (nvec << 4) + threadIdx.x; i < rowBytes; i += blockDim.x) dst[i] = src[i];
    } else {
      for (uint32_t i = threadIdx.x; i < rowBytes; i += blockDim.x) dst[i] = src[i];
    }
  }
}

/* FP32 FMA ceiling probe: 8 independent dependent-chains per thread, registers only. */
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
      x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace rtg
#endif
