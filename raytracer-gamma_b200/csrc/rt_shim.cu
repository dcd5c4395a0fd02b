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
#include "rt_soa.h"

using namespace rtg;

#define RT_DEFAULT_MIN_BLOCKS 2
#ifndef RT_DEFAULT_SLOTS
#define RT_DEFAULT_SLOTS 4
#endif
#define RT_LOCKSTEP_CTA_MAX 384u      /* lockstep passes of the whole CTA up to this many filter records (rt_kernels.cuh trace_body) */
#define RT_LOCKSTEP_MAX_SPHERES 896u  /* ... of half-CTAs up to this many; free-running warps beyond */
#define RT_COPY_CHUNK (8u << 20)     /* pageable readback: D2H and the host memcpy alternate over two pinned chunks */
#define RT_FLUSH_BYTES (256u << 20)  /* rt_cuda_flush_l2: larger than the 126 MB L2 */

struct rt_cuda_ctx {
  int device = 0;
  int smCount = 0;
  int smemOptin = 0;               /* largest dynamic shared memory one CTA may ask for */
  int smemPerSM = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  char lastError[256] = {0};

  /* scene */
  bool haveScene = false, haveClusters = false;
  uint32_t n = 0, nPad = 0, nl = 0;
  float4_* dScene = nullptr;       /* one allocation: filt | geo | matA | matB | lpos | lcol | cluster form */
  size_t sceneCap = 0;
  SceneView view{};
  SceneLayout layout{};
  std::vector<rt_sphere> hSpheres; /* kept for the lazy build of the cluster form (option accel) */
  std::vector<rt_light> hLights;
  ConstRecords* hConst = nullptr;  /* padded filter records, handed to trace_kernel_const as a launch parameter */

  /* frame */
  bool haveFrame = false;
  uint32_t W = 0, H = 0, localRows = 0;
  float4* dFb = nullptr;  size_t fbCap = 0;        /* pixels */
  float* dPacked = nullptr; size_t packedCap = 0;  /* floats */
  float4* dSamples = nullptr; size_t samplesCap = 0;   /* per-sample results when spp > 1 */
  unsigned char* dRgb8 = nullptr; size_t rgbCap = 0;
  unsigned int* dWork = nullptr;     /* [0] queue head, [1] max bits */
  unsigned long long* dCounters = nullptr;
  Frame* dFrames = nullptr; size_t framesCap = 0;       /* the threads' stacks of suspended calls (64 B per frame) */
  unsigned int* dOrder = nullptr; size_t orderCap = 0;   /* work-order state of a launch (rt_kernels.cuh "Work order") */
  void* dFlush = nullptr;

  /* synchronous readback through two pinned chunks */
  unsigned char* hChunk[2] = {nullptr, nullptr};
  cudaEvent_t evChunk[2] = {nullptr, nullptr};

  /* asynchronous readback: two frames in flight on a second stream */
  cudaStream_t copyStream = nullptr;
  struct Ticket {
    unsigned char* dRgb = nullptr; size_t cap = 0;
    cudaEvent_t evQuant = nullptr, evCopy = nullptr;
    bool pending = false;
    unsigned char* hStage = nullptr; size_t stageCap = 0;   /* used when the caller's buffer is not pinned */
    unsigned char* userDst = nullptr; size_t bytes = 0;
  } tk[2];
  unsigned nextTicket = 0;

  /* options */
  int staging = 0, noFilter = 0, blocksPerSM = 0, slots = 0;
  int order = 0;                   /* 0 auto (= 1) | 1: deep tiles first | 2: tiles in scanline order */
  int sweepStep = 0, deepAt = 0;   /* work-order tuning (0 = default) */
  int lockstep = 0;                /* 0 auto (by the number of filter records) | 1 whole CTA | 2 off | 3 half-CTAs */
  int slotMode = 0;                /* 0 auto | 1 slot records in shared memory | 2 in local memory */
  int accel = 0;                   /* 1: two-level cluster filter (optional accelerated mode) */
  uint32_t nc = 0, ncPad = 0;

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
    case RT_CUDA_ERR_NCCL: return "RT_CUDA_ERR_NCCL";
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
  ctx->smemPerSM = (int)pr.sharedMemPerMultiprocessor;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(RT_CUDA_ERR_CUDA);
  ctx->ownStream = true;
  if (cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess) return fail(RT_CUDA_ERR_CUDA);
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) return fail(RT_CUDA_ERR_CUDA);
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreateWithFlags(&ctx->evChunk[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->tk[i].evQuant, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->tk[i].evCopy, cudaEventDisableTiming) != cudaSuccess)
      return fail(RT_CUDA_ERR_CUDA);
  }
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
  if (ctx->copyStream) cudaStreamSynchronize(ctx->copyStream);
  cudaFree(ctx->dScene); cudaFree(ctx->dFb); cudaFree(ctx->dPacked); cudaFree(ctx->dRgb8); cudaFree(ctx->dSamples);
  cudaFree(ctx->dWork); cudaFree(ctx->dCounters); cudaFree(ctx->dFlush); cudaFree(ctx->dOrder); cudaFree(ctx->dFrames);
  for (int i = 0; i < 2; ++i) {
    if (ctx->hChunk[i]) cudaFreeHost(ctx->hChunk[i]);
    if (ctx->evChunk[i]) cudaEventDestroy(ctx->evChunk[i]);
    cudaFree(ctx->tk[i].dRgb);
    if (ctx->tk[i].hStage) cudaFreeHost(ctx->tk[i].hStage);
    if (ctx->tk[i].evQuant) cudaEventDestroy(ctx->tk[i].evQuant);
    if (ctx->tk[i].evCopy) cudaEventDestroy(ctx->tk[i].evCopy);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx->hConst;
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

extern "C" void* rt_cuda_get_stream(rt_cuda_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int rt_cuda_get_device(rt_cuda_ctx* ctx) { return ctx ? ctx->device : -1; }

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
  if (!strcmp(key, "blocks_per_sm")) { if (value < 0 || value > 8) return RT_CUDA_ERR_INVALID_ARG; ctx->blocksPerSM = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "order")) { if (value < 0 || value > 2) return RT_CUDA_ERR_INVALID_ARG; ctx->order = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "lockstep")) { if (value < 0 || value > 3) return RT_CUDA_ERR_INVALID_ARG; ctx->lockstep = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "sweep_step")) { if (value < 0 || value > 32) return RT_CUDA_ERR_INVALID_ARG; ctx->sweepStep = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "deep_at")) { if (value < 0 || value > RT_COUNT_MAX) return RT_CUDA_ERR_INVALID_ARG; ctx->deepAt = (int)value; return RT_CUDA_OK; }
  if (!strcmp(key, "slot_mode")) { if (value < 0 || value > 2) return RT_CUDA_ERR_INVALID_ARG; ctx->slotMode = (int)value; return RT_CUDA_OK; }
#ifdef RT_DEV_VARIANTS   /* development builds only: slot-count variants of the local-memory kernel */
  if (!strcmp(key, "slots")) { if (value != 0 && (value < 3 || value > 4)) return RT_CUDA_ERR_INVALID_ARG; ctx->slots = (int)value; return RT_CUDA_OK; }
#endif
  return RT_CUDA_ERR_INVALID_ARG;
}

/* (Re)build the device scene from the context's host copy; `withClusters` adds the two-level form. */
static int push_scene(rt_cuda_ctx* ctx, bool withClusters) {
  std::vector<float4_> h;
  SceneLayout lay;
  build_scene_soa(ctx->hSpheres.data(), (uint32_t)ctx->hSpheres.size(), ctx->hLights.data(),
                  (uint32_t)ctx->hLights.size(), h, lay, withClusters);
  if (lay.total > ctx->sceneCap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dScene); ctx->dScene = nullptr; ctx->sceneCap = 0;
    CU(cudaMalloc(&ctx->dScene, lay.total * sizeof(float4_)));
    ctx->sceneCap = lay.total;
  }
  CU(cudaMemcpyAsync(ctx->dScene, h.data(), lay.total * sizeof(float4_), cudaMemcpyHostToDevice, ctx->stream));
  if (lay.n && lay.nPad <= RT_CONST_MAX_SPHERES) {
    if (!ctx->hConst) ctx->hConst = new (std::nothrow) ConstRecords();
    if (!ctx->hConst) return RT_CUDA_ERR_OUT_OF_MEMORY;
    memcpy(ctx->hConst->r, h.data() + lay.offFilt, (size_t)lay.nPad * sizeof(float4_));
    for (uint32_t i = lay.nPad; i < RT_CONST_MAX_SPHERES; ++i) ctx->hConst->r[i] = float4_{0.f, 0.f, 0.f, INFINITY};
  }
  CU(cudaStreamSynchronize(ctx->stream));   /* h goes out of scope */
  ctx->layout = lay;
  ctx->view = scene_view(ctx->dScene, lay);
  ctx->n = lay.n; ctx->nPad = lay.nPad; ctx->nl = lay.nl; ctx->nc = lay.nc; ctx->ncPad = lay.ncPad;
  ctx->haveClusters = withClusters;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_upload_scene(rt_cuda_ctx* ctx, const rt_sphere* spheres, unsigned sphNum,
                                    const rt_light* lights, unsigned lgtNum) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  if ((sphNum && !spheres) || (lgtNum && !lights)) return RT_CUDA_ERR_INVALID_ARG;
  if (sphNum > RT_CUDA_MAX_SPHERES || lgtNum > RT_CUDA_MAX_LIGHTS) return RT_CUDA_ERR_TOO_LARGE;
  CU(cudaSetDevice(ctx->device));
  ctx->haveScene = false;
  ctx->hSpheres.assign(spheres, spheres + sphNum);
  ctx->hLights.assign(lights, lights + lgtNum);
  /* the cluster form is only built when the accelerated mode is (or later becomes) selected */
  int rc = push_scene(ctx, ctx->accel != 0);
  if (rc) return rc;
  ctx->haveScene = true;
  return RT_CUDA_OK;
}

/* grow-only device buffer */
template <typename T>
static int ensure_dev(rt_cuda_ctx* ctx, T*& ptr, size_t& cap, size_t need) {
  if (need > cap) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ptr); ptr = nullptr; cap = 0;
    CU(cudaMalloc(&ptr, need * sizeof(T)));
    cap = need;
  }
  return RT_CUDA_OK;
}

#define RT_ACCEL_MIN_SPHERES 768u           /* accel = 1 engages from here; accel = 2 forces it */

extern "C" int rt_cuda_render_strips(rt_cuda_ctx* ctx, unsigned width, unsigned height, float zoom,
                                     float aliasFactor, int maxStack, unsigned stripRows,
                                     unsigned stripFirst, unsigned stripStride) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveScene) return RT_CUDA_ERR_NO_SCENE;
  if (width == 0 || height == 0 || maxStack < 1 || maxStack > RT_CUDA_MAX_STACK) return RT_CUDA_ERR_INVALID_ARG;
  if (stripRows == 0 || stripStride == 0 || stripFirst >= stripStride) return RT_CUDA_ERR_INVALID_ARG;
  if ((uint64_t)width * height >= (1ull << 31)) return RT_CUDA_ERR_TOO_LARGE;
  CU(cudaSetDevice(ctx->device));

  /* ---- everything that can fail is computed and allocated before the context is touched ---- */
  const uint32_t nStrips = (height + stripRows - 1) / stripRows;
  uint32_t localRows = 0;
  for (uint32_t s = stripFirst; s < nStrips; s += stripStride) {
    const uint32_t r0 = s * stripRows;
    const uint32_t r1 = (r0 + stripRows < height) ? r0 + stripRows : height;
    localRows += r1 - r0;
  }
  const size_t pixels = (size_t)localRows * width;

  TraceParams p;
  p.cam = make_camera(width, height, zoom, aliasFactor, maxStack, (int)ctx->n);
  if (p.cam.nIter >= (1 << 20)) return RT_CUDA_ERR_TOO_LARGE;          /* make_camera's cap: alias too large to iterate */
  const uint64_t spp64 = (uint64_t)p.cam.nIter * (uint64_t)p.cam.nIter;
  if (spp64 >= (1ull << 32)) return RT_CUDA_ERR_TOO_LARGE;
  const uint32_t spp = (uint32_t)spp64;
  p.wm = make_workmap(width, localRows, stripRows, stripFirst, stripStride, spp, (uint32_t)p.cam.nIter);
  if ((uint64_t)p.wm.tilesX * p.wm.tilesY * 32u * p.wm.K >= (1ull << 32)) return RT_CUDA_ERR_TOO_LARGE;
  /* work order: every tile's first group from the queue head, then the tiles marked deep, then a sweep over the
   * rest (rt_kernels.cuh "Work order") */
  bool lpt = ctx->order != 2 && p.wm.K > 1u;

  /* the accelerated mode needs something to cull, its records in shared memory, and the cluster form */
  const size_t listBytes = (size_t)RT_LIST_MAX * RT_BLOCK * sizeof(unsigned short);
  const uint32_t list1Max = (ctx->n >= 2048u) ? RT_LIST1_LARGE : RT_LIST1_SMALL;
  const bool wantAccel = ctx->accel && !ctx->noFilter && ctx->n >= (ctx->accel == 2 ? 4u * RT_CLUSTER : RT_ACCEL_MIN_SPHERES);
  if (wantAccel && !ctx->haveClusters) {       /* option set after the upload: add the cluster form now */
    int rc = push_scene(ctx, true);
    if (rc) return rc;
  }
  const size_t accelBytes = (size_t)ctx->ncPad * (16 + RT_CLUSTER * 18) + (size_t)list1Max * RT_BLOCK * sizeof(unsigned short);
  /* ... and falls back to the plain mode when they do not fit one CTA (> ~8 000 spheres) */
  const bool accel = wantAccel && ctx->haveClusters && 16 + listBytes + accelBytes <= (size_t)ctx->smemOptin;

  /* staging: shared memory filled by TMA bulk copies is the default at every size (it is equal or faster
   * than the constant bank from 16 spheres up, profiles/r1/sweep_config5.jsonl); option staging=1 selects
   * the constant-bank path for scenes of <= RT_CONST_MAX_SPHERES records */
  int staging = ctx->staging;
  if (staging == 0) staging = 2;
  if (staging == 1 && (ctx->n == 0 || ctx->nPad > RT_CONST_MAX_SPHERES || !ctx->hConst)) staging = 2;
  if (accel) staging = 2;
  const bool useConst = (staging == 1);
  const size_t sceneBytes = useConst ? 0 : accel ? (size_t)ctx->ncPad * (16 + RT_CLUSTER * 18) : (size_t)ctx->nPad * 16;
  const size_t smemBase = 16 + sceneBytes + (accel ? (size_t)list1Max * RT_BLOCK * sizeof(unsigned short) : 0) + listBytes;
  /* slot records on chip whenever two CTAs per SM still fit: four slots per lane, else three; else local memory */
  const size_t slotBytes1 = (size_t)RT_SLOT_WORDS * RT_BLOCK * sizeof(uint32_t);
  const size_t perCtaMax = (size_t)ctx->smemPerSM / RT_DEFAULT_MIN_BLOCKS - 1024;      /* 1 KB per CTA is reserved by the system */
  int smemSlots = 0;                 /* 0 = local memory, else the slot count */
  if (!accel && ctx->slotMode != 2) {
    for (int ns = RT_SLOTS; ns >= 3 && !smemSlots; --ns) {
      const size_t need = smemBase + slotBytes1 * ns;
      if (need <= (size_t)ctx->smemOptin && (need <= perCtaMax || (ctx->slotMode == 1 && ns == 3))) smemSlots = ns;
    }
#ifdef RT_DEV_VARIANTS
    if (ctx->slots == 3 && smemSlots == 4) smemSlots = 3;
#endif
  }
  const size_t smem = smemBase + slotBytes1 * smemSlots;

  void (*kern)(const TraceParams) = nullptr;
  void (*kernC)(const TraceParams, const ConstRecords) = nullptr;
  if (accel) kern = trace_kernel<RT_DEFAULT_MIN_BLOCKS, RT_SLOTS, true, false>;
  else if (useConst) kernC = smemSlots == 4 ? trace_kernel_const<RT_DEFAULT_MIN_BLOCKS, 4, true>
                           : smemSlots == 3 ? trace_kernel_const<RT_DEFAULT_MIN_BLOCKS, 3, true>
                                            : trace_kernel_const<RT_DEFAULT_MIN_BLOCKS, RT_SLOTS, false>;
  else kern = smemSlots == 4 ? trace_kernel<RT_DEFAULT_MIN_BLOCKS, 4, false, true>
            : smemSlots == 3 ? trace_kernel<RT_DEFAULT_MIN_BLOCKS, 3, false, true>
                             : trace_kernel<RT_DEFAULT_MIN_BLOCKS, RT_SLOTS, false, false>;
  const void* kfn = kernC ? (const void*)kernC : (const void*)kern;
  int perSM = 0;
  uint32_t grid = 0, totalWork = 0;
  if (pixels && spp) {
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kfn, RT_BLOCK, smem));
    if (perSM < 1) return RT_CUDA_ERR_TOO_LARGE;
    if (ctx->blocksPerSM > 0 && ctx->blocksPerSM < perSM) perSM = ctx->blocksPerSM;
    totalWork = p.wm.nTiles * 32u * p.wm.K;
    grid = (uint32_t)ctx->smCount * (uint32_t)perSM;
    const uint32_t needBlocks = (totalWork + RT_BLOCK - 1) / RT_BLOCK;
    if (grid > needBlocks) grid = needBlocks;
    if (grid < 1) grid = 1;
  }
  /* ... which needs the first groups to last long enough for deep samples to show: several times the samples in flight
   * (measured: a 68-row shard of the 8K frame is 7 % slower with it, a 540-row shard 4.6 % faster) */
  if (ctx->order == 0 && (uint64_t)p.wm.nTiles * 32u < 4ull * grid * RT_BLOCK * RT_SLOTS) lpt = false;
  int rc = ensure_dev(ctx, ctx->dFb, ctx->fbCap, pixels ? pixels : 1);
  if (rc) { ctx->haveFrame = false; return rc; }
  if (pixels && spp) {
    rc = ensure_dev(ctx, ctx->dFrames, ctx->framesCap, (size_t)grid * RT_BLOCK * RT_SLOTS * (size_t)maxStack);
    if (rc) { ctx->haveFrame = false; return rc; }
  }
  const size_t orderWords = 32 + 2 * (size_t)p.wm.nTiles;
  if (lpt && pixels) {
    rc = ensure_dev(ctx, ctx->dOrder, ctx->orderCap, orderWords);
    if (rc) { ctx->haveFrame = false; return rc; }
  }
  if (spp > 1 && pixels) {
    rc = ensure_dev(ctx, ctx->dSamples, ctx->samplesCap, pixels * spp);
    if (rc) { ctx->haveFrame = false; return rc; }
  }

  /* ---- publish and launch ---- */
  ctx->haveFrame = false;
  ctx->launches = 0;
  ctx->timed = false;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  CU(cudaMemsetAsync(ctx->dWork, 0, 4 * sizeof(unsigned int), ctx->stream));
  CU(cudaMemsetAsync(ctx->dCounters, 0, RT_NUM_COUNTERS * sizeof(unsigned long long), ctx->stream));
  ctx->W = width; ctx->H = height; ctx->localRows = localRows;
  ctx->stats.engine = 1;
  if (pixels == 0) { ctx->haveFrame = true; return RT_CUDA_OK; }
  if (spp == 0) {          /* alias <= 0: the sample loops never run, the frame is black (main.cpp:420) */
    CU(cudaMemsetAsync(ctx->dFb, 0, pixels * sizeof(float4), ctx->stream));
    ctx->haveFrame = true;
    return RT_CUDA_OK;
  }

  p.sc = ctx->view;
  p.fb = ctx->dFb;
  p.workCounter = ctx->dWork;
  p.maxBits = ctx->dWork + 1;
  p.counters = ctx->dCounters;
  p.lpt = lpt ? 1u : 0u;
  p.total1 = lpt ? p.wm.nTiles * 32u : totalWork;
  /* tiles per sweep claim: ONE.  Larger claims (even guided ones that shrink towards the end) measured slower — 16.2 vs
   * 14.4 ms at 256 spheres / 4K with up to 8: neighbouring tiles are equally deep, and a claim of several lands all of
   * them on one warp */
  p.sweepStep = ctx->sweepStep ? (uint32_t)ctx->sweepStep : 1u;
  /* lockstep passes where the per-pass loop is short (the accelerated mode loops over clusters): the whole CTA up to
   * RT_LOCKSTEP_CTA_MAX records, halves of it up to RT_LOCKSTEP_MAX_SPHERES (measured at 4K, profiles/r2/ab_lockstep_groups.txt:
   * 256 records 13.8 ms with 8 warps, 14.2 with 4, 18.5 free; 512: 26.0 / 24.9 / 30.1; 768: 42.3 / 41.4 / 43.5; 1 024: 201.7 / 193.5 / 192.0) */
  const uint32_t ctaMax = accel ? 64u : RT_LOCKSTEP_CTA_MAX, halfMax = accel ? 512u : RT_LOCKSTEP_MAX_SPHERES;   /* accelerated mode: cluster records
      (1 024 spheres = 128 clusters: 40.5 ms with 8 warps, 38.2 with 4, 48.9 free; 256 spheres = 32 clusters: 15.1 / 16.4 / 22.1) */
  const uint32_t loopRecords = accel ? ctx->ncPad : ctx->nPad;
  p.lockstep = (ctx->lockstep == 1) ? 8u : (ctx->lockstep == 3) ? 4u : (ctx->lockstep == 2) ? 0u
             : (loopRecords <= ctaMax) ? 8u : (loopRecords <= halfMax) ? 4u : 0u;
  p.frames = ctx->dFrames;
  p.deepAt = ctx->deepAt ? (uint32_t)ctx->deepAt : RT_DEEP_AT;
  p.tileClaimed = p.deepList = p.orderCtl = nullptr;
  if (lpt) {
    /* one zeroed allocation: control words (32) | tileClaimed[nTiles] | deepList[nTiles] */
    CU(cudaMemsetAsync(ctx->dOrder, 0, orderWords * sizeof(unsigned int), ctx->stream));
    p.orderCtl = ctx->dOrder;
    p.tileClaimed = ctx->dOrder + 32;
    p.deepList = p.tileClaimed + p.wm.nTiles;
  }
  p.spp = spp;
  p.samples = (spp > 1) ? ctx->dSamples : nullptr;
  p.noFilter = ctx->noFilter;
  p.list1Max = list1Max;
  /* queue granule: small enough to balance the tail, large enough to amortise the atomic */
  const uint32_t warps = grid * (RT_BLOCK / 32);
  uint32_t chunk = (p.total1 / (warps * 16u)) & ~31u;
  if (chunk < 32u) chunk = 32u;
  if (chunk > 256u || lpt) chunk = 256u;     /* the first groups need no fine balance: the deep list and the sweep follow */
  p.chunk = chunk;

  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  if (kernC) kernC<<<grid, RT_BLOCK, smem, ctx->stream>>>(p, *ctx->hConst);
  else       kern<<<grid, RT_BLOCK, smem, ctx->stream>>>(p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += 1;
  if (spp > 1) {
    combine_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(ctx->dSamples, ctx->dFb, (uint32_t)pixels, spp, ctx->dWork + 1);
    CU(cudaGetLastError());
    ctx->launches += 1;
  }
  ctx->stats.accel = accel ? 1u : 0u;
  ctx->stats.clusters = accel ? ctx->nc : 0u;
  ctx->stats.grid = grid; ctx->stats.block = RT_BLOCK; ctx->stats.smem_bytes = (uint32_t)smem;
  ctx->stats.staging = (uint32_t)staging;
  ctx->stats.slots_on_chip = (uint32_t)smemSlots;
  ctx->haveFrame = true;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_render(rt_cuda_ctx* ctx, unsigned width, unsigned height, float zoom,
                              float aliasFactor, int maxStack) {
  return rt_cuda_render_strips(ctx, width, height, zoom, aliasFactor, maxStack,
                               height ? height : 1, 0, 1);
}

extern "C" int rt_cuda_pack(rt_cuda_ctx* ctx) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  CU(cudaSetDevice(ctx->device));
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  if (pixels == 0) return RT_CUDA_OK;
  int rc = ensure_dev(ctx, ctx->dPacked, ctx->packedCap, pixels * 3);
  if (rc) return rc;
  const int grid = ctx->smCount * 8;
  pack_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->dFb, ctx->dPacked, (uint32_t)pixels);
  CU(cudaGetLastError());
  ctx->launches += 1;
  return RT_CUDA_OK;
}

static int quantise_into(rt_cuda_ctx* ctx, void* devDst, float maxColour, cudaStream_t stream) {
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  const int grid = ctx->smCount * 8;
  quantise_kernel<<<grid, 256, 0, stream>>>(ctx->dFb, reinterpret_cast<uint32_t*>(devDst),
                                            (uint32_t)pixels, ctx->dWork + 1, maxColour);
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
  int rc = ensure_dev(ctx, ctx->dRgb8, ctx->rgbCap, ((pixels + 3) / 4) * 12);
  if (rc) return rc;
  return quantise_into(ctx, ctx->dRgb8, maxColour, ctx->stream);
}

extern "C" int rt_cuda_quantise_to(rt_cuda_ctx* ctx, void* devDst, size_t dstBytes, float maxColour) {
  if (!ctx || !devDst) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  if (pixels == 0) return RT_CUDA_OK;
  /* the kernel stores whole 32-bit words when four pixels are complete and single bytes for the rest */
  if (dstBytes < pixels * 3 || ((uintptr_t)devDst & 3u)) return RT_CUDA_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  return quantise_into(ctx, devDst, maxColour, ctx->stream);
}

extern "C" void* rt_cuda_device_packed(rt_cuda_ctx* ctx) { return ctx ? (void*)ctx->dPacked : nullptr; }
extern "C" void* rt_cuda_device_rgb8(rt_cuda_ctx* ctx) { return ctx ? (void*)ctx->dRgb8 : nullptr; }
extern "C" void* rt_cuda_device_max(rt_cuda_ctx* ctx) { return ctx ? (void*)(ctx->dWork + 1) : nullptr; }

extern "C" void* rt_cuda_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void rt_cuda_host_free(void* p) { if (p) { cudaFreeHost(p); cudaGetLastError(); } }

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

/* Device -> caller memory.  A pinned destination (rt_cuda_host_alloc, cudaHostRegister) takes one DMA;
 * pageable memory goes through two pinned chunks so the DMA of chunk i+1 runs while the CPU copies chunk i. */
static int copy_out(rt_cuda_ctx* ctx, void* dst, const void* dsrc, size_t bytes) {
  if (is_pinned(dst)) {
    CU(cudaMemcpyAsync(dst, dsrc, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return RT_CUDA_OK;
  }
  for (int i = 0; i < 2; ++i)
    if (!ctx->hChunk[i]) CU(cudaMallocHost(&ctx->hChunk[i], RT_COPY_CHUNK));
  const size_t nChunks = (bytes + RT_COPY_CHUNK - 1) / RT_COPY_CHUNK;
  for (size_t c = 0; c <= nChunks; ++c) {
    if (c < nChunks) {
      const size_t off = c * (size_t)RT_COPY_CHUNK, n = (bytes - off < RT_COPY_CHUNK) ? bytes - off : RT_COPY_CHUNK;
      CU(cudaMemcpyAsync(ctx->hChunk[c & 1], (const unsigned char*)dsrc + off, n, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaEventRecord(ctx->evChunk[c & 1], ctx->stream));
    }
    if (c > 0) {
      const size_t off = (c - 1) * (size_t)RT_COPY_CHUNK, n = (bytes - off < RT_COPY_CHUNK) ? bytes - off : RT_COPY_CHUNK;
      CU(cudaEventSynchronize(ctx->evChunk[(c - 1) & 1]));
      memcpy((unsigned char*)dst + off, ctx->hChunk[(c - 1) & 1], n);
    }
  }
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

/* Asynchronous readback: quantise on the render stream into one of two device buffers, copy on a
 * second stream, return at once.  The next rt_cuda_render overlaps the copy. */
extern "C" int rt_cuda_readback_rgb8_async(rt_cuda_ctx* ctx, unsigned char* dst, float maxColour, int* ticket) {
  if (!ctx || !dst || !ticket) return RT_CUDA_ERR_INVALID_ARG;
  if (!ctx->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  CU(cudaSetDevice(ctx->device));
  const unsigned t = ctx->nextTicket & 1u;
  rt_cuda_ctx::Ticket& k = ctx->tk[t];
  if (k.pending) {            /* a third frame: finish the oldest first */
    int rc = rt_cuda_readback_wait(ctx, (int)t);
    if (rc) return rc;
  }
  const size_t pixels = (size_t)ctx->localRows * ctx->W;
  const size_t bytes = pixels * 3;
  *ticket = (int)t;
  ctx->nextTicket++;
  k.userDst = nullptr; k.bytes = 0;
  if (!pixels) { k.pending = false; return RT_CUDA_OK; }
  int rc = ensure_dev(ctx, k.dRgb, k.cap, ((pixels + 3) / 4) * 12);
  if (rc) return rc;
  unsigned char* target = dst;
  if (!is_pinned(dst)) {      /* pageable destination: stage in pinned memory, copied out by the wait */
    if (bytes > k.stageCap) {
      if (k.hStage) cudaFreeHost(k.hStage);
      k.hStage = nullptr; k.stageCap = 0;
      CU(cudaMallocHost(&k.hStage, bytes));
      k.stageCap = bytes;
    }
    target = k.hStage;
    k.userDst = dst; k.bytes = bytes;
  }
  rc = quantise_into(ctx, k.dRgb, maxColour, ctx->stream);
  if (rc) return rc;
  CU(cudaEventRecord(k.evQuant, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->copyStream, k.evQuant, 0));
  CU(cudaMemcpyAsync(target, k.dRgb, bytes, cudaMemcpyDeviceToHost, ctx->copyStream));
  CU(cudaEventRecord(k.evCopy, ctx->copyStream));
  k.pending = true;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_readback_wait(rt_cuda_ctx* ctx, int ticket) {
  if (!ctx || ticket < 0 || ticket > 1) return RT_CUDA_ERR_INVALID_ARG;
  rt_cuda_ctx::Ticket& k = ctx->tk[ticket];
  if (!k.pending) return RT_CUDA_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventSynchronize(k.evCopy));
  if (k.userDst) memcpy(k.userDst, k.hStage, k.bytes);
  k.pending = false; k.userDst = nullptr;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_flush_l2(rt_cuda_ctx* ctx) {
  if (!ctx) return RT_CUDA_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->dFlush) CU(cudaMalloc(&ctx->dFlush, RT_FLUSH_BYTES));
  CU(cudaMemsetAsync(ctx->dFlush, (int)(ctx->launches & 0xFF), RT_FLUSH_BYTES, ctx->stream));
  return RT_CUDA_OK;
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
