/* rt_shim.cu — the C-ABI of include/rt_cuda.h over the kernels in rt_kernels.cuh.
 * Replaces the OpenCL host plumbing of main.cpp:182-350,456-489 (see rt_cuda.h
 * for the call-by-call mapping).  No CPU rendering path exists in this file. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "rt_cuda.h"
#include "rt_kernels.cuh"
#include "rt_wavefront.cuh"
#include "rt_soa.h"

using namespace rtg;

#define RT_DEFAULT_MIN_BLOCKS 2
#define RT_DEFAULT_SLOTS 4

struct rt_cuda_ctx {
  int device = 0;
  int smCount = 0;
  int smemOptin = 0;               /* largest dynamic shared memory one CTA may ask for */
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  char lastError[256] = {0};

  /* scene */
  bool haveScene = false;
  uint32_t n = 0, nPad = 0, nl = 0;
  float4_* dScene = nullptr;       /* one allocation: filt | geo | matA | matB | lpos | lcol */
  size_t sceneCap = 0;
  SceneView view{};

  /* frame */
  bool haveFrame = false;
  uint32_t W = 0, H = 0, localRows = 0;
  float4* dFb = nullptr;  size_t fbCap = 0;        /* pixels */
  float* dPacked = nullptr; size_t packedCap = 0;  /* floats */
  float4* dSamples = nullptr; size_t samplesCap = 0;   /* per-sample results when spp > 1 */
  unsigned char* dRgb8 = nullptr; size_t rgbCap = 0;
  unsigned int* dWork = nullptr;     /* [0] queue head, [1] max bits */
  unsigned long long* dCounters = nullptr;
  void* hPinned = nullptr; size_t pinnedCap = 0;

  /* options */
  int staging = 0, noFilter = 0, blocksPerSM = 0, minBlocks = 0, slots = 0, prefetch = 0;
  int engine = 0;                  /* 0 auto, 1 persistent kernel, 2 wavefront */
  int accel = 0;                   /* 1: two-level cluster filter (optional accelerated mode) */
  uint32_t nc = 0, ncPad = 0;
  long pool = 0;                   /* wavefront: samples in flight (0 = default) */

  /* wavefront pool */
  uint32_t* dPool = nullptr; size_t poolCap = 0;      /* words */
  Frame* dStacks = nullptr; size_t stacksCap = 0;     /* samples */
  uint32_t* dQueues = nullptr; size_t queuesCap = 0;  /* entries */
  uint32_t* dCounts = nullptr;                        /* [2][4] */
  uint32_t* hCounts = nullptr;                        /* pinned [4] */

  /* stats */
  rt_cuda_stats stats{};
  uint32_t launches = 0;
};

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      snprintf(ctx->lastError, sizeof ctx->lastError, "%s: %s", #call,                   \
               cudaGetErrorString(e_));                                                  \
      return (e_ == cudaErrorMemoryAllocation) ? RT_CUDA_ERR_OUT_OF_MEMORY               \
                                               : RT_CUDA_ERR_CUDA;                       \
    }                                                                                    \
  } while (0)

extern "C" const char* rt_cuda_strerror(int status) {
  switch (status) {
    case RT_CUDA_OK: return "RT_CUDA_OK";
    case RT_CUDA_ERR_INVALID_ARG: return "RT_CUDA_ERR_INVALID_ARG";
    case RT_CUDA_ERR_NO_DEVICE: return "RT_CUDA_ERR_NO_DEVICE";
    case RT_CUDA_ERR_CUDA: return "RT_CUDA_ERR_CUDA";
    case RT_CUDA_ERR_NO_SCENE: return "RT_CUDA_ERR_NO_SCENE";
    case RT_CUDA_ERR_NO_FRAME: return "RT_CUDA_ERR_NO_FRAME";
    case RT_CUDA_ERR_TOO_LARGE: return "RT_CUDA_ERR_TOO_LARGE";
    case RT_CUDA_ERR_OUT_OF_MEMORY: return "RT_CUDA_ERR_OUT_OF_MEMORY";
    default: return "RT_CUDA_ERR_UNKNOWN";
  }
}

extern "C" const char* rt_cuda_last_error(rt_cuda_ctx* ctx) { return ctx ? ctx->lastError : ""; }

extern "C" int rt_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int rt_cuda_device_info(int device, char* buf, size_t bufSize) {
  if (!buf || bufSize == 0) return RT_CUDA_ERR_INVALID_ARG;
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) { cudaGetLastError(); return RT_CUDA_ERR_NO_DEVICE; }
  int clockKHz = 0;
  cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, device);
  snprintf(buf, bufSize,
           "Device is %s  compute %d.%d  %d SMs  max clock %d MHz  %.1f GB global  "
           "%zu KB shared/SM  %d regs/SM  max block %d",
           pr.name, pr.major, pr.minor, pr.multiProcessorCount, clockKHz / 1000,
           (double)pr.totalGlobalMem / 1e9, pr.sharedMemPerMultiprocessor / 1024,
           pr.regsPerMultiprocessor, pr.maxThreadsPerBlock);
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_init(int device, rt_cuda_ctx** out) {
  if (!out) return RT_CUDA_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return RT_CUDA_ERR_NO_DEVICE; }
  if (device < 0 || device >= count) return RT_CUDA_ERR_INVALID_ARG;
  rt_cuda_ctx* ctx = new (std::nothrow) rt_cuda_ctx();
  if (!ctx) return RT_CUDA_ERR_OUT_OF_MEMORY;
  ctx->device = device;
  auto fail = [&](int code) { rt_cuda_destroy(ctx); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return fail(RT_CUDA_ERR_CUDA);
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) return fail(RT_CUDA_ERR_CUDA);
  if (pr.major < 10) {   /* built for sm_100a only */
    return fail(RT_CUDA_ERR_NO_DEVICE);
  }
  ctx->smCount = pr.multiProcessorCount;
  ctx->smemOptin = (int)pr.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(RT_CUDA_ERR_CUDA);
  ctx->ownStream = true;
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) return fail(RT_CUDA_ERR_CUDA);
  if (cudaMalloc(&ctx->dWork, 4 * sizeof(unsigned int)) != cudaSuccess) return fail(RT_CUDA_ERR_OUT_OF_MEMORY);
  if (cudaMalloc(&ctx->dCounters, RT_NUM_COUNTERS * sizeof(unsigned long long)) != cudaSuccess) return fail(RT_CUDA_ERR_OUT_OF_MEMORY);
  cudaMemset(ctx->dWork, 0, 4 * sizeof(unsigned int));
  cudaMemset(ctx->dCounters, 0, RT_NUM_COUNTERS * sizeof(unsigned long long));
  *out = ctx;
  return RT_CUDA_OK;
}

extern "C" void rt_cuda_destroy(rt_cuda_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->dScene); cudaFree(ctx->dFb); cudaFree(ctx->dPacked); cudaFree(ctx->dRgb8); cudaFree(ctx->dSamples);
  cudaFree(ctx->dPool); cudaFree(ctx->dStacks); cudaFree(ctx->dQueues); cudaFree(ctx->dCounts);
  if (ctx->hCounts) cudaFreeHost(ctx->hCounts);
  cudaFree(ctx->dWork); cudaFree(ctx->dCounters);
  if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
}

extern "C" int rt_cuda_set_stream(rt_cuda_ctx* ctx, void* s) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  if (ctx->stream) CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->ownStream = false;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_synchronize(rt_cuda_ctx* ctx) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_set_option(rt_cuda_ctx* ctx, const char* key, long value) {
  if (!ctx || !key) return RT_CUDA_ERR_INVALID_ARG;
  if (!strcmp(key, "staging")) { if (value < 0 || value > 2) return RT_CUDA_ERR_INVALID_ARG; ctx->staging = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "no_filter")) { ctx->noFilter = value ? 1 : 0; return RT_CUDA_OK; }
  if (!strcmp(key, "accel")) { if (value < 0 || value > 2) return RT_CUDA_ERR_INVALID_ARG; ctx->accel = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "engine")) { if (value < 0 || value > 2) return RT_CUDA_ERR_INVALID_ARG; ctx->engine = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "pool")) { if (value < 0) return RT_CUDA_ERR_INVALID_ARG; ctx->pool = value; return RT_CUDA_OK; }
  if (!strcmp(key, "prefetch")) { ctx->prefetch = value ? 1 : 0; return RT_CUDA_OK; }
  if (!strcmp(key, "slots")) { if (value != 0 && (value < 2 || value > 4)) return RT_CUDA_ERR_INVALID_ARG; ctx->slots = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "min_blocks")) { if (value != 0 && (value < 2 || value > 4)) return RT_CUDA_ERR_INVALID_ARG; ctx->minBlocks = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "blocks_per_sm")) { if (value < 0 || value > 8) return RT_CUDA_ERR_INVALID_ARG; ctx->blocksPerSM = (int)value; return RT_CUDA_OK; }
  return RT_CUDA_ERR_INVALID_ARG;
}

extern "C" int rt_cuda_upload_scene(rt_cuda_ctx* ctx, const rt_sphere* spheres, unsigned sphNum,
                                    const rt_light* lights, unsigned lgtNum) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  if ((sphNum && !spheres) || (lgtNum && !lights)) return RT_CUDA_ERR_INVALID_ARG;
  if (sphNum > RT_CUDA_MAX_SPHERES) return RT_CUDA_ERR_TOO_LARGE;
  CU(cudaSetDevice(ctx->device));
  std::vector<float4_> h;
  SceneLayout lay;
  build_scene_soa(spheres, sphNum, lights, lgtNum, h, lay);

  if (lay.total > ctx->sceneCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dScene); ctx->dScene = nullptr; ctx->sceneCap = 0;
    CU(cudaMalloc(&ctx->dScene, lay.total * sizeof(float4_)));
    ctx->sceneCap = lay.total;
  }
  CU(cudaMemcpyAsync(ctx->dScene, h.data(), lay.total * sizeof(float4_), cudaMemcpyHostToDevice, ctx->stream));
  if (lay.n && lay.nPad <= RT_CONST_MAX_SPHERES)
    CU(cudaMemcpyToSymbolAsync(c_filt, h.data() + lay.offFilt, (size_t)lay.nPad * sizeof(float4_), 0,
                               cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));   /* h goes out of scope */

  ctx->view = scene_view(ctx->dScene, lay);
  ctx->n = lay.n; ctx->nPad = lay.nPad; ctx->nl = lay.nl; ctx->nc = lay.nc; ctx->ncPad = lay.ncPad;
  ctx->haveScene = true;
  return RT_CUDA_OK;
}

static int ensure_frame(rt_cuda_ctx* ctx, size_t pixels) {
  if (pixels > ctx->fbCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dFb); ctx->dFb = nullptr; ctx->fbCap = 0;
    CU(cudaMalloc(&ctx->dFb, pixels * sizeof(float4)));
    ctx->fbCap = pixels;
  }
  return RT_CUDA_OK;
}

#define RT_ACCEL_MIN_SPHERES 768u           /* accel = 1 engages from here (6 % slower at 512, 25 % faster at 1024: profiles/r1/sweep_config5.jsonl); accel = 2 forces it */
#define RT_DEFAULT_POOL (1l << 21)          /* samples in flight of the wavefront engine */

/* The wavefront engine: alternate the pure filter kernels and the shade kernel over a pool
 * of samples until the frame's work counter and all queues are drained (rt_wavefront.cuh).
 * The host only launches; it looks at the queue lengths every few rounds to know when to stop. */
static int render_wavefront(rt_cuda_ctx* ctx, const TraceParams& tp, uint32_t spp, size_t pixels) {
  const uint64_t totalWork = (uint64_t)pixels * spp;
  if (totalWork >= (1ull << 32)) return RT_CUDA_ERR_TOO_LARGE;
  uint64_t P = (uint64_t)(ctx->pool > 0 ? ctx->pool : RT_DEFAULT_POOL);
  if (P > totalWork) P = totalWork;
  if (P * RT_SLOT_WORDS > ctx->poolCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dPool); ctx->dPool = nullptr; ctx->poolCap = 0;
    CU(cudaMalloc(&ctx->dPool, P * RT_SLOT_WORDS * sizeof(uint32_t)));
    ctx->poolCap = P * RT_SLOT_WORDS;
  }
  if (P > ctx->stacksCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dStacks); ctx->dStacks = nullptr; ctx->stacksCap = 0;
    CU(cudaMalloc(&ctx->dStacks, P * RT_MAX_STACK * sizeof(Frame)));
    ctx->stacksCap = P;
  }
  if (P * 6 > ctx->queuesCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dQueues); ctx->dQueues = nullptr; ctx->queuesCap = 0;
    CU(cudaMalloc(&ctx->dQueues, P * 6 * sizeof(uint32_t)));
    ctx->queuesCap = P * 6;
  }
  if (!ctx->dCounts) CU(cudaMalloc(&ctx->dCounts, 8 * sizeof(uint32_t)));
  if (!ctx->hCounts) CU(cudaMallocHost(&ctx->hCounts, 4 * sizeof(uint32_t)));

  WfParams wp;
  wp.sc = tp.sc; wp.cam = tp.cam;
  wp.st = ctx->dPool; wp.stacks = ctx->dStacks; wp.queues = ctx->dQueues; wp.counts = ctx->dCounts;
  wp.P = (uint32_t)P; wp.cur = 0;
  wp.fb = tp.fb; wp.samples = tp.samples; wp.spp = spp;
  wp.workCounter = tp.workCounter; wp.maxBits = tp.maxBits; wp.counters = tp.counters;
  wp.localRows = tp.localRows; wp.stripRows = tp.stripRows; wp.stripFirst = tp.stripFirst; wp.stripStride = tp.stripStride;
  wp.totalWork = (uint32_t)totalWork;
  wp.noFilter = tp.noFilter;

  const size_t smem = 16 + (size_t)ctx->nPad * 16 + (size_t)RT_LIST_MAX * RT_BLOCK * sizeof(unsigned short)
                      + (size_t)4 * RT_SHADOW_BATCH * RT_BLOCK * sizeof(float);
  CU(cudaFuncSetAttribute(wf_filter<K_TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(cudaFuncSetAttribute(wf_filter<K_SHADOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(cudaFuncSetAttribute(wf_filter<K_CONTAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSM = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, wf_filter<K_SHADOW>, RT_BLOCK, smem));
  if (perSM < 1) return RT_CUDA_ERR_TOO_LARGE;
  const uint32_t gridF = (uint32_t)ctx->smCount * (uint32_t)perSM;
  const uint32_t gridS = (uint32_t)ctx->smCount * 16u;

  /* the first P work items are handed out by wf_spawn; the work counter continues from P */
  const uint32_t init[8] = {(uint32_t)P, 0, 0, 0, 0, 0, 0, 0};
  CU(cudaMemcpyAsync(ctx->dCounts, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  const uint32_t first = (uint32_t)P;
  CU(cudaMemcpyAsync(tp.workCounter, &first, sizeof first, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));          /* `init` / `first` are stack variables */

  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  wf_spawn<<<(uint32_t)((P + 255) / 256), 256, 0, ctx->stream>>>(wp);
  CU(cudaGetLastError());
  ctx->launches += 1;
  const int CHECK = 4;
  for (int round = 1;; ++round) {
    wf_filter<K_TRACE><<<gridF, RT_BLOCK, smem, ctx->stream>>>(wp);
    wf_filter<K_SHADOW><<<gridF, RT_BLOCK, smem, ctx->stream>>>(wp);
    wf_filter<K_CONTAIN><<<gridF, RT_BLOCK, smem, ctx->stream>>>(wp);
    CU(cudaMemsetAsync(ctx->dCounts + (wp.cur ^ 1u) * 4u, 0, 4 * sizeof(uint32_t), ctx->stream));
    wf_shade<<<gridS, 128, 0, ctx->stream>>>(wp, K_TRACE);
    wf_shade<<<gridS, 128, 0, ctx->stream>>>(wp, K_SHADOW);
    wf_shade<<<gridS, 128, 0, ctx->stream>>>(wp, K_CONTAIN);
    CU(cudaGetLastError());
    ctx->launches += 6;
    wp.cur ^= 1u;
    if (round % CHECK == 0) {
      CU(cudaMemcpyAsync(ctx->hCounts, ctx->dCounts + wp.cur * 4u, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      if ((ctx->hCounts[0] | ctx->hCounts[1] | ctx->hCounts[2]) == 0u) break;
    }
    if (round > (1 << 22)) return RT_CUDA_ERR_CUDA;   /* cannot happen: every round retires work */
  }
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->stats.grid = gridF; ctx->stats.block = RT_BLOCK; ctx->stats.smem_bytes = (uint32_t)smem;
  ctx->stats.staging = 2;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_render_strips(rt_cuda_ctx* ctx, unsigned width, unsigned height, float zoom,
                                     float aliasFactor, int maxStack, unsigned stripRows,
                                     unsigned stripFirst, unsigned stripStride) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveScene) return RT_CUDA_ERR_NO_SCENE;
  if (width == 0 || height == 0 || maxStack < 1 || maxStack > RT_CUDA_MAX_STACK) return RT_CUDA_ERR_INVALID_ARG;
  if (stripRows == 0 || stripStride == 0 || stripFirst >= stripStride) return RT_CUDA_ERR_INVALID_ARG;
  if ((uint64_t)width * height >= (1ull << 31)) return RT_CUDA_ERR_TOO_LARGE;
  CU(cudaSetDevice(ctx->device));

  /* rows owned by this shard */
  const uint32_t nStrips = (height + stripRows - 1) / stripRows;
  uint32_t localRows = 0;
  for (uint32_t s = stripFirst; s < nStrips; s += stripStride) {
    const uint32_t r0 = s * stripRows;
    const uint32_t r1 = (r0 + stripRows < height) ? r0 + stripRows : height;
    localRows += r1 - r0;
  }
  ctx->W = width; ctx->H = height; ctx->localRows = localRows;
  ctx->launches = 0;
  ctx->timed = false;
  const size_t pixels = (size_t)localRows * width;
  int rc = ensure_frame(ctx, pixels ? pixels : 1);
  if (rc) return rc;

  CU(cudaMemsetAsync(ctx->dWork, 0, 4 * sizeof(unsigned int), ctx->stream));
  CU(cudaMemsetAsync(ctx->dCounters, 0, RT_NUM_COUNTERS * sizeof(unsigned long long), ctx->stream));
  ctx->haveFrame = true;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  if (pixels == 0) return RT_CUDA_OK;

  TraceParams p;
  p.sc = ctx->view;
  p.cam = make_camera(width, height, zoom, aliasFactor, maxStack, (int)ctx->n);
  p.fb = ctx->dFb;
  p.workCounter = ctx->dWork;
  p.maxBits = ctx->dWork + 1;
  p.counters = ctx->dCounters;
  p.localRows = localRows;
  p.stripRows = stripRows; p.stripFirst = stripFirst; p.stripStride = stripStride;
  /* work item = one sample of one pixel; samples are summed in order by combine_kernel */
  const uint32_t spp = (uint32_t)p.cam.nIter * (uint32_t)p.cam.nIter;
  if (spp == 0) {          /* alias <= 0: the sample loops never run, the frame is black (main.cpp:420) */
    CU(cudaMemsetAsync(ctx->dFb, 0, pixels * sizeof(float4), ctx->stream));
    return RT_CUDA_OK;
  }
  p.tilesX = (width + 7u) / 8u;
  const uint32_t tilesY = (localRows + 3u) / 4u;
  if ((uint64_t)p.tilesX * tilesY * 32u * spp >= (1ull << 32)) return RT_CUDA_ERR_TOO_LARGE;
  p.totalWork = p.tilesX * tilesY * 32u * spp;
  p.spp = spp;
  p.samples = nullptr;
  if (spp > 1) {
    const size_t need = pixels * spp;
    if (need > ctx->samplesCap) {
      CU(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->dSamples); ctx->dSamples = nullptr; ctx->samplesCap = 0;
      CU(cudaMalloc(&ctx->dSamples, need * sizeof(float4)));
      ctx->samplesCap = need;
    }
    p.samples = ctx->dSamples;
  }
  p.noFilter = ctx->noFilter;
  p.prefetch = ctx->prefetch;

  /* engine choice: the wavefront engine (separate filter / shade kernels) is opt-in */
  int engine = ctx->engine;
  if (engine == 0) engine = 1;    /* the persistent kernel is faster at every size measured */
  if (engine == 2 && ctx->n > 0) {
    int rc = render_wavefront(ctx, p, spp, pixels);
    if (rc) return rc;
    ctx->stats.engine = 2;
    if (spp > 1) {
      combine_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(ctx->dSamples, ctx->dFb, (uint32_t)pixels, spp, ctx->dWork + 1);
      CU(cudaGetLastError());
      ctx->launches += 1;
    }
    return RT_CUDA_OK;
  }
  ctx->stats.engine = 1;

  /* staging choice: __constant__ broadcast for small scenes, shared memory (TMA bulk) otherwise */
  int staging = ctx->staging;
  if (staging == 0) staging = (ctx->n > 0 && ctx->nPad <= 64) ? 1 : 2;
  if (staging == 1 && (ctx->n == 0 || ctx->nPad > RT_CONST_MAX_SPHERES)) staging = 2;
  /* the accelerated mode needs something to cull and its records in shared memory */
  const size_t perCta = (size_t)RT_LIST_MAX * RT_BLOCK * sizeof(unsigned short) + (size_t)4 * RT_SHADOW_BATCH * RT_BLOCK * sizeof(float) + 16;
  const uint32_t list1Max = (ctx->n >= 2048u) ? RT_LIST1_LARGE : RT_LIST1_SMALL;
  const size_t accelBytes = (size_t)ctx->ncPad * (16 + RT_CLUSTER * 18) + (size_t)list1Max * RT_BLOCK * sizeof(unsigned short);
  /* ... and falls back to the plain mode when they do not fit one CTA (> ~8 000 spheres) */
  const bool accel = ctx->accel && !ctx->noFilter && ctx->n >= (ctx->accel == 2 ? 4u * RT_CLUSTER : RT_ACCEL_MIN_SPHERES) &&
                     perCta + accelBytes <= (size_t)ctx->smemOptin;
  if (accel) staging = 2;
  const bool useConst = (staging == 1);
  const size_t sceneBytes = useConst ? 0 : accel ? (size_t)ctx->ncPad * (16 + RT_CLUSTER * 18) : (size_t)ctx->nPad * 16;
  p.list1Max = list1Max;
  const size_t smem = 16 + sceneBytes + (accel ? (size_t)list1Max * RT_BLOCK * sizeof(unsigned short) : 0)
                      + (size_t)RT_LIST_MAX * RT_BLOCK * sizeof(unsigned short)
                      + (size_t)4 * RT_SHADOW_BATCH * RT_BLOCK * sizeof(float);
  /* variants: MIN_BLOCKS resident CTAs per SM (register budget), NSLOTS pixels in flight per lane */
  const int minBlocks = ctx->minBlocks ? ctx->minBlocks : RT_DEFAULT_MIN_BLOCKS;
  const int nslots = ctx->slots ? ctx->slots : RT_DEFAULT_SLOTS;
  void (*kern)(const TraceParams) = nullptr;
#define RT_PICK(C, M) ((nslots == 2) ? trace_kernel<C, M, 2> : (nslots == 3) ? trace_kernel<C, M, 3> : trace_kernel<C, M, 4>)
  if (useConst) kern = (minBlocks == 2) ? RT_PICK(true, 2) : (minBlocks == 3) ? RT_PICK(true, 3) : RT_PICK(true, 4);
  else          kern = (minBlocks == 2) ? RT_PICK(false, 2) : (minBlocks == 3) ? RT_PICK(false, 3) : RT_PICK(false, 4);
#undef RT_PICK
  if (accel) kern = trace_kernel<false, RT_DEFAULT_MIN_BLOCKS, RT_DEFAULT_SLOTS, true>;
  ctx->stats.accel = accel ? 1u : 0u;
  ctx->stats.clusters = accel ? ctx->nc : 0u;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSM = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, RT_BLOCK, smem));
  if (perSM < 1) return RT_CUDA_ERR_TOO_LARGE;
  if (ctx->blocksPerSM > 0 && ctx->blocksPerSM < perSM) perSM = ctx->blocksPerSM;
  uint32_t grid = (uint32_t)ctx->smCount * (uint32_t)perSM;
  const uint32_t needBlocks = (p.totalWork + RT_BLOCK - 1) / RT_BLOCK;
  if (grid > needBlocks) grid = needBlocks;
  if (grid < 1) grid = 1;
  /* queue granule: small enough to balance the tail, large enough to amortise the atomic */
  const uint32_t warps = grid * (RT_BLOCK / 32);
  uint32_t chunk = (p.totalWork / (warps * 16u)) & ~31u;
  if (chunk < 32u) chunk = 32u;
  if (chunk > 256u) chunk = 256u;
  p.chunk = chunk;

  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  kern<<<grid, RT_BLOCK, smem, ctx->stream>>>(p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += 1;
  if (spp > 1) {
    combine_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(ctx->dSamples, ctx->dFb, (uint32_t)pixels, spp, ctx->dWork + 1);
    CU(cudaGetLastError());
    ctx->launches += 1;
  }
  ctx->stats.grid = grid; ctx->stats.block = RT_BLOCK; ctx->stats.smem_bytes = (uint32_t)smem;
  ctx->stats.staging = (uint32_t)staging;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_render(rt_cuda_ctx* ctx, unsigned width, unsigned height, float zoom,
                              float aliasFactor, int maxStack) {
  return rt_cuda_render_strips(ctx, width, height, zoom, aliasFactor, maxStack,
                               height ? height : 1, 0, 1);
}

static int ensure_pinned(rt_cuda_ctx* ctx, size_t bytes) {
  if (bytes > ctx->pinnedCap) {
    if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
    ctx->hPinned = nullptr; ctx->pinnedCap = 0;
    CU(cudaMallocHost(&ctx->hPinned, bytes));
    ctx->pinnedCap = bytes;
  }
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_pack(rt_cuda_ctx* ctx) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  CU(cudaSetDevice(ctx->device));
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  if (pixels == 0) return RT_CUDA_OK;
  if (pixels * 3 > ctx->packedCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dPacked); ctx->dPacked = nullptr; ctx->packedCap = 0;
    CU(cudaMalloc(&ctx->dPacked, pixels * 3 * sizeof(float)));
    ctx->packedCap = pixels * 3;
  }
  const int grid = ctx->smCount * 8;
  pack_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->dFb, ctx->dPacked, (uint32_t)pixels);
  CU(cudaGetLastError());
  ctx->launches += 1;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_quantise(rt_cuda_ctx* ctx, float maxColour) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  CU(cudaSetDevice(ctx->device));
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  if (pixels == 0) return RT_CUDA_OK;
  const size_t bytes = ((pixels + 3) / 4) * 12;
  if (bytes > ctx->rgbCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dRgb8); ctx->dRgb8 = nullptr; ctx->rgbCap = 0;
    CU(cudaMalloc(&ctx->dRgb8, bytes));
    ctx->rgbCap = bytes;
  }
  const int grid = ctx->smCount * 8;
  quantise_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->dFb, reinterpret_cast<uint32_t*>(ctx->dRgb8),
                                                 (uint32_t)pixels, ctx->dWork + 1, maxColour);
  CU(cudaGetLastError());
  ctx->launches += 1;
  return RT_CUDA_OK;
}

extern "C" void* rt_cuda_device_packed(rt_cuda_ctx* ctx) { return ctx ? (void*)ctx->dPacked : nullptr; }
extern "C" void* rt_cuda_device_rgb8(rt_cuda_ctx* ctx) { return ctx ? (void*)ctx->dRgb8 : nullptr; }
extern "C" void* rt_cuda_device_max(rt_cuda_ctx* ctx) { return ctx ? (void*)(ctx->dWork + 1) : nullptr; }

/* Copy through the context's pinned staging buffer so the transfer runs at full PCIe rate
 * whatever memory the caller handed us. */
static int copy_out(rt_cuda_ctx* ctx, void* dst, const void* dsrc, size_t bytes) {
  int rc = ensure_pinned(ctx, bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->hPinned, dsrc, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(dst, ctx->hPinned, bytes);
  return RT_CUDA_OK;
}

static int fetch_max(rt_cuda_ctx* ctx, float* outMax) {
  unsigned bits = 0;
  CU(cudaMemcpyAsync(&bits, ctx->dWork + 1, sizeof bits, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  float m; memcpy(&m, &bits, sizeof m);
  ctx->stats.max_colour = m;
  if (outMax) *outMax = (m == 0.f) ? 1.f : m;   /* algebra.h:86-88 */
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_readback(rt_cuda_ctx* ctx, rt_vec* dst, float* outMax) {
  if (!ctx || !dst) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  int rc = rt_cuda_pack(ctx);
  if (rc) return rc;
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  if (pixels) {
    rc = copy_out(ctx, dst, ctx->dPacked, pixels * 3 * sizeof(float));
    if (rc) return rc;
  }
  return fetch_max(ctx, outMax);
}

extern "C" int rt_cuda_readback_rgb8(rt_cuda_ctx* ctx, unsigned char* dst, float maxColour) {
  if (!ctx || !dst) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  int rc = rt_cuda_quantise(ctx, maxColour);
  if (rc) return rc;
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  if (!pixels) return RT_CUDA_OK;
  return copy_out(ctx, dst, ctx->dRgb8, pixels * 3);
}

extern "C" int rt_cuda_get_stats(rt_cuda_ctx* ctx, rt_cuda_stats* out) {
  if (!ctx || !out) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  unsigned long long c[RT_NUM_COUNTERS];
  CU(cudaMemcpy(c, ctx->dCounters, sizeof c, cudaMemcpyDeviceToHost));
  rt_cuda_stats& s = ctx->stats;
  s.rays = c[0]; s.shadow_rays = c[1]; s.contain_queries = c[2]; s.contain_tests = c[3];
  s.exact_tests = c[4]; s.samples = c[5]; s.null_rays = c[6];
  s.lane_iters = c[10];                         /* sub-query capacity of all sphere passes    */
  s.active_lane_iters = c[7] + c[8] + c[9];     /* sub-queries actually served (trace/shadow/contain) */
  s.served_trace = c[7]; s.served_shadow = c[8]; s.served_contain = c[9]; s.passes = c[11];
  s.passes_trace = c[12]; s.passes_shadow2 = c[13]; s.passes_shadow4 = c[14]; s.passes_contain = c[15];
  for (int i = 0; i < 6; ++i) s.phase_cycles[i] = c[16 + i];
  s.filter_tests = ctx->noFilter ? 0 : c[10] * (unsigned long long)(s.accel ? ctx->ncPad : ctx->nPad);   /* accelerated mode: cluster tests only */
  s.sph_num = ctx->n; s.sph_padded = ctx->nPad; s.lgt_num = ctx->nl;
  s.width = ctx->W; s.height = ctx->H; s.local_rows = ctx->localRows;
  s.kernel_ms = 0.f;
  if (ctx->timed) CU(cudaEventElapsedTime(&s.kernel_ms, ctx->ev0, ctx->ev1));
  int rc = fetch_max(ctx, nullptr);
  if (rc) return rc;
  s.kernel_launches = ctx->launches;
  *out = s;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_assemble_rgb8(rt_cuda_ctx* ctx, const void* gathered, void* out, unsigned width,
                                     unsigned height, unsigned stripRows, unsigned nShards,
                                     size_t shardPitchBytes) {
  if (!ctx || !gathered || !out || !width || !height || !stripRows || !nShards) return RT_CUDA_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  const unsigned grid = height < (unsigned)ctx->smCount * 8u ? height : (unsigned)ctx->smCount * 8u;
  assemble_rgb8_kernel<<<grid, 256, 0, ctx->stream>>>((const unsigned char*)gathered, (unsigned char*)out,
                                                      width, height, stripRows, nShards, shardPitchBytes);
  CU(cudaGetLastError());
  ctx->launches += 1;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_ffma_peak(rt_cuda_ctx* ctx, int iters, float* outTflops) {
  if (!ctx || !outTflops || iters < 1) return RT_CUDA_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  const int blocks = ctx->smCount * 8, threads = 256;
  float* d = nullptr;
  CU(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  ffma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters / 8 + 1, 0.999f, 0.001f);   /* warm-up */
  CU(cudaEventRecord(e0, ctx->stream));
  ffma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999f, 0.001f);
  CU(cudaEventRecord(e1, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  const double flops = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
  *outTflops = (float)(flops / (ms * 1e-3) / 1e12);
  ctx->launches += 2;
  return RT_CUDA_OK;
}
