/* rt_kernels.cuh — sm_100a kernels of the trace loop.
 *
 * trace_kernel     persistent CTAs, one lane = one pixel in flight.  Replaces
 *                  `__kernel raytrace` (raytrace_kernel.cl:870-973) and the CPU
 *                  pixel loop (main.cpp:404-453).
 * pack_kernel      float4 framebuffer -> packed Vec[W*H] (the reference's dst layout)
 * quantise_kernel  float4 framebuffer + max -> RGB8 (main.cpp:71-76)
 *
 * Loop structure of trace_kernel (per warp, all lanes converged at every step):
 *   refill   lanes without a pixel take the next work items from a tile queue
 *            (global atomicAdd per warp chunk, __ballot_sync ranks the takers)
 *   filter   every lane runs its current query (ray or containment probe)
 *            against ALL spheres: 1 broadcast LDS.128 + 7 FFMA + 1 FSETP + 1 mask
 *            op per sphere, sphere records staged once per CTA into shared memory
 *            by a TMA bulk copy (cp.async.bulk + mbarrier) or read from
 *            __constant__ for small scenes
 *   gather   lanes whose 32-sphere mask is non-zero append candidate indices to a
 *            per-lane list in shared memory
 *   resolve  k-th candidate of every lane is put through the reference's exact
 *            expressions together (convergent)
 *   advance  O(1) shading / state transition (divergent by query kind, short)
 */
#ifndef RT_KERNELS_CUH
#define RT_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "rt_core.cuh"

namespace rtg {

#define RT_BLOCK 256
#define RT_LIST_MAX 8
#define RT_CONST_MAX_SPHERES 1024
#define RT_NUM_COUNTERS 10

struct TraceParams {
  SceneView sc;
  Camera cam;
  float4* fb;               /* [localRows*W] {r,g,b,1}                              */
  unsigned int* workCounter;/* tile queue head                                      */
  unsigned int* maxBits;    /* running max of positive channel values (float bits)  */
  unsigned long long* counters;  /* [RT_NUM_COUNTERS] rays, shadow, containQ, containT, exact, samples, lane-iters, active-lane-iters, null rays */
  uint32_t localRows;       /* rows rendered by this context                        */
  uint32_t stripRows, stripFirst, stripStride;   /* row r is ours iff (r/stripRows)%stripStride==stripFirst */
  uint32_t tilesX, totalWork, chunk;
  int noFilter;             /* debug: exact test for every sphere                   */
};

__constant__ float4_ c_filt[RT_CONST_MAX_SPHERES];

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

/* Map a queue index to a pixel: 8x4 tiles, row-major inside, tiles row-major. */
__device__ __forceinline__ bool work_to_pixel(const TraceParams& p, uint32_t idx, uint32_t& gx,
                                              uint32_t& gy, uint32_t& dst) {
  const uint32_t tile = idx >> 5, within = idx & 31u;
  const uint32_t ty = tile / p.tilesX, tx = tile - ty * p.tilesX;
  const uint32_t x = tx * 8u + (within & 7u);
  const uint32_t y = ty * 4u + (within >> 3);
  if (x >= p.cam.W || y >= p.localRows) return false;
  const uint32_t strip = y / p.stripRows;
  gx = x;
  gy = (strip * p.stripStride + p.stripFirst) * p.stripRows + (y - strip * p.stripRows);
  dst = y * p.cam.W + x;
  return true;
}

/* List overflow (more than RT_LIST_MAX candidates for one query): resolve what is queued.
 * Rare and divergent, so it is kept out of line (and by value, so the lane state stays in
 * registers) to keep the filter loop small. */
__device__ __noinline__ HitAcc flush_list(HitAcc h, int kind, V3 o, V3 d, const float4_* geo,
                                          uint32_t n, const unsigned short* sList, uint32_t tid) {
#pragma unroll 1
  for (int k = 0; k < RT_LIST_MAX; ++k) {
    const uint32_t i = sList[k * RT_BLOCK + tid];
    if (i < n) h = resolve_one(h, kind, o, d, geo[i], i);
  }
  return h;
}

template <bool USE_CONST, int MIN_BLOCKS, int GROUP>
__global__ void __launch_bounds__(RT_BLOCK, MIN_BLOCKS) trace_kernel(const TraceParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  /* layout: [mbarrier 16 B][filter records nPad*16 B (shared staging only)][lists] */
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float4* sFilt = reinterpret_cast<float4*>(smem_raw + 16);
  const uint32_t filtBytes = USE_CONST ? 0u : p.sc.nPad * 16u;
  unsigned short* sList = reinterpret_cast<unsigned short*>(smem_raw + 16 + filtBytes);

  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const unsigned FULL = 0xFFFFFFFFu;

  if (!USE_CONST) {
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, filtBytes);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(p.sc.filt);
      unsigned char* dstp = reinterpret_cast<unsigned char*>(sFilt);
      for (uint32_t off = 0; off < filtBytes; off += 32768u) {
        const uint32_t n = (filtBytes - off < 32768u) ? (filtBytes - off) : 32768u;
        tma_bulk_g2s(dstp + off, src + off, n, bar);
      }
    }
    mbar_wait(bar, 0);
  }

  Lane L;
  Frame stack[RT_MAX_STACK];
  L.kind = K_NULL;
  query_null(L.qy);
  L.pixel = 0xFFFFFFFFu;
  L.minT = 1000.f; L.hitIdx = -1;
  L.top = -1;
  L.ctr.rays = L.ctr.shadow = L.ctr.containQ = L.ctr.containT = L.ctr.exactTests = L.ctr.samples = 0;
  L.ctr.nullRays = 0;
  L.medium = (int)p.sc.n; L.obj = 0; L.light = 0;
  float laneMax = 0.f;
  uint32_t laneIters = 0, activeIters = 0;

  uint32_t wbase = 0, wend = 0;   /* warp-uniform slice of the tile queue */
  bool queueDry = false;

  const uint32_t groups = p.sc.nPad / GROUP;   /* nPad is a multiple of 32 */

  for (;;) {
    /* ---- refill ---- */
    bool need = (L.pixel == 0xFFFFFFFFu);
    while (!queueDry) {
      const unsigned m = __ballot_sync(FULL, need);
      if (m == 0) break;
      if (wbase >= wend) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(p.workCounter, p.chunk);
        b = __shfl_sync(FULL, b, 0);
        if (b >= p.totalWork) { queueDry = true; break; }
        wbase = b;
        wend = (b + p.chunk < p.totalWork) ? b + p.chunk : p.totalWork;
      }
      const uint32_t avail = wend - wbase;
      const uint32_t rank = __popc(m & ((1u << lane) - 1u));
      if (need && rank < avail) {
        uint32_t gx, gy, dst;
        if (work_to_pixel(p, wbase + rank, gx, gy, dst)) {
          if (start_pixel(L, p.cam, gx, gy, dst)) {
            need = false;
          } else {
            /* zero samples: the pixel is black (main.cpp:420, loop never runs) */
            p.fb[dst] = make_float4(0.f, 0.f, 0.f, 1.f);
            L.pixel = 0xFFFFFFFFu;
          }
        }
      }
      const uint32_t cnt = __popc(m);
      wbase += (cnt < avail) ? cnt : avail;
    }
    if (__ballot_sync(FULL, L.kind != K_NULL) == 0) break;

    /* ---- filter + gather ---- */
    int cnt = 0;
    laneIters++;
    if (L.kind != K_NULL) activeIters++;
    const bool filterable = query_filterable(L.qy);
    if (!p.noFilter) {
      for (uint32_t g = 0; g < groups; ++g) {
        /* bit (GROUP-1-j) of `skip` = sign bit of test j: one FADD + one funnel shift per test */
        unsigned skip = 0;
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {
          float4 s4;
          if (USE_CONST) {
            const float4_ c = c_filt[g * GROUP + j];
            s4 = make_float4(c.x, c.y, c.z, c.w);
          } else {
            s4 = sFilt[g * GROUP + j];
          }
          float4_ s; s.x = s4.x; s.y = s4.y; s.z = s4.z; s.w = s4.w;
          skip = __funnelshift_l(__float_as_uint(filter_diff(L.qy, s)), skip, 1);
        }
        unsigned cand = ~skip & (GROUP == 32 ? 0xFFFFFFFFu : ((1u << (GROUP & 31)) - 1u));
        if (cand) {
          do {
            const int b = 31 - __clz(cand);          /* highest bit = lowest sphere index */
            cand &= ~(1u << b);
            if (cnt == RT_LIST_MAX) {
              HitAcc h; h.minT = L.minT; h.hitIdx = L.hitIdx;
              h = flush_list(h, L.kind, query_origin(L.qy), mk(L.qy.dx, L.qy.dy, L.qy.dz), p.sc.geo,
                             p.sc.n, sList, tid);
              L.minT = h.minT; L.hitIdx = h.hitIdx;
              L.ctr.exactTests += RT_LIST_MAX;
              cnt = 0;
            }
            sList[cnt * RT_BLOCK + tid] = (unsigned short)(g * GROUP + (GROUP - 1 - b));
            ++cnt;
          } while (cand);
        }
      }
      /* ---- resolve (convergent over lanes) ---- */
      if (filterable) {
        const int maxc = __reduce_max_sync(FULL, cnt);
#pragma unroll 1
        for (int k = 0; k < maxc; ++k) {
          if (k < cnt) resolve_candidate(L, p.sc, sList[k * RT_BLOCK + tid]);
        }
      } else {
        /* not reached by any finite scene; keeps NaN/inf geometry on the reference's own path */
        (void)__reduce_max_sync(FULL, 0);
        L.minT = 1000.f; L.hitIdx = -1;
        if (L.kind != K_NULL)
          for (uint32_t i = 0; i < p.sc.n; ++i) resolve_candidate(L, p.sc, i);
      }
    } else {
      if (L.kind != K_NULL && L.qy.q != INFINITY)
        for (uint32_t i = 0; i < p.sc.n; ++i) resolve_candidate(L, p.sc, i);
    }

    /* ---- advance ---- */
    if (L.kind != K_NULL) {
      if (advance(L, p.sc, stack, p.cam)) {
        if (finish_sample(L, p.cam)) {
          p.fb[L.pixel] = make_float4(L.acc.x, L.acc.y, L.acc.z, 1.f);
          if (L.acc.x > laneMax) laneMax = L.acc.x;   /* algebra.h:74-82, NaN skipped */
          if (L.acc.y > laneMax) laneMax = L.acc.y;
          if (L.acc.z > laneMax) laneMax = L.acc.z;
          L.pixel = 0xFFFFFFFFu;
          L.kind = K_NULL;
          query_null(L.qy);
        }
      }
    }
  }

  /* ---- per-warp reductions ---- */
  unsigned mb = __float_as_uint(laneMax);   /* laneMax >= 0: uint order == float order */
  mb = __reduce_max_sync(FULL, mb);
  const unsigned r0 = __reduce_add_sync(FULL, L.ctr.rays);
  const unsigned r1 = __reduce_add_sync(FULL, L.ctr.shadow);
  const unsigned r2 = __reduce_add_sync(FULL, L.ctr.containQ);
  unsigned long long ct = L.ctr.containT;
  for (int o = 16; o > 0; o >>= 1) ct += __shfl_xor_sync(FULL, ct, o);
  unsigned long long ex = L.ctr.exactTests;
  for (int o = 16; o > 0; o >>= 1) ex += __shfl_xor_sync(FULL, ex, o);
  const unsigned r5 = __reduce_add_sync(FULL, L.ctr.samples);
  const unsigned r8 = __reduce_add_sync(FULL, L.ctr.nullRays);
  unsigned long long li = laneIters, ai = activeIters;
  for (int o = 16; o > 0; o >>= 1) { li += __shfl_xor_sync(FULL, li, o); ai += __shfl_xor_sync(FULL, ai, o); }
  if (lane == 0) {
    if (mb) atomicMax(p.maxBits, mb);
    atomicAdd(&p.counters[0], (unsigned long long)r0);
    atomicAdd(&p.counters[1], (unsigned long long)r1);
    atomicAdd(&p.counters[2], (unsigned long long)r2);
    atomicAdd(&p.counters[3], ct);
    atomicAdd(&p.counters[4], ex);
    atomicAdd(&p.counters[5], (unsigned long long)r5);
    atomicAdd(&p.counters[6], li);
    atomicAdd(&p.counters[7], ai);
    atomicAdd(&p.counters[8], (unsigned long long)r8);
  }
}

/* float4 framebuffer -> packed 12-byte pixels (the reference's `Vec dst[]`, .cl:972) */
__global__ void pack_kernel(const float4* __restrict__ fb, float* __restrict__ out, uint32_t npix) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t nflt = npix * 3u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nflt; i += stride) {
    const uint32_t px = i / 3u, c = i - px * 3u;
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
