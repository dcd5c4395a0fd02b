/* rt_multi.cpp — the multi-GPU C-ABI of include/rt_cuda_multi.h (librt_cuda_multi.so).
 *
 * Sequences the single-GPU C-ABI (rt_cuda.h) and NCCL; no kernels of its own.  Replaces the
 * reference's single cl_command_queue (main.cpp:229) by one stream per GPU of the box, with
 * the one exchange a frame needs (global maximum, algebra.h:68-91, and the gather of the
 * quantised strips).  See the header for the frame sequence. */
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <vector>

#include "rt_cuda_multi.h"

static_assert(sizeof(ncclUniqueId) == RT_CUDA_MULTI_ID_BYTES, "ncclUniqueId size");

struct Local {
  rt_cuda_ctx* ctx = nullptr;
  int device = 0;
  int rank = 0;
  ncclComm_t comm = nullptr;
  unsigned char* dGather = nullptr; size_t gatherCap = 0;   /* world blocks of `pitch` bytes */
  unsigned char* dFrame = nullptr;  size_t frameCap = 0;    /* H*W*3 */
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t copyStream = nullptr;                        /* asynchronous readback of dFrame */
  cudaEvent_t evFrame = nullptr, evCopied = nullptr;
  bool copyPending = false;
};

struct rt_cuda_multi {
  int world = 0;
  std::vector<Local> L;
  char lastError[256] = {0};
  bool haveFrame = false;
  unsigned W = 0, H = 0;
};

#define MCU(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      snprintf(m->lastError, sizeof m->lastError, "%s: %s", #call, cudaGetErrorString(e_));    \
      return (e_ == cudaErrorMemoryAllocation) ? RT_CUDA_ERR_OUT_OF_MEMORY : RT_CUDA_ERR_CUDA; \
    }                                                                                          \
  } while (0)
#define MNC(call)                                                                              \
  do {                                                                                         \
    ncclResult_t r_ = (call);                                                                  \
    if (r_ != ncclSuccess) {                                                                   \
      snprintf(m->lastError, sizeof m->lastError, "%s: %s", #call, ncclGetErrorString(r_));    \
      return RT_CUDA_ERR_NCCL;                                                                 \
    }                                                                                          \
  } while (0)
#define MRT(l, call)                                                                           \
  do {                                                                                         \
    int s_ = (call);                                                                           \
    if (s_ != RT_CUDA_OK) {                                                                    \
      snprintf(m->lastError, sizeof m->lastError, "%s (rank %d): %s %s", #call, (l).rank,      \
               rt_cuda_strerror(s_), rt_cuda_last_error((l).ctx));                             \
      return s_;                                                                               \
    }                                                                                          \
  } while (0)

/* ---- row-strip arithmetic (host only; the same arithmetic as assemble_rgb8_kernel) ---- */
extern "C" unsigned rt_cuda_multi_shard_rows(unsigned height, unsigned stripRows, unsigned rank, unsigned world) {
  if (!stripRows || !world || rank >= world) return 0;
  const unsigned nStrips = (height + stripRows - 1) / stripRows;
  unsigned rows = 0;
  for (unsigned s = rank; s < nStrips; s += world) {
    const unsigned r0 = s * stripRows;
    rows += ((r0 + stripRows < height) ? r0 + stripRows : height) - r0;
  }
  return rows;
}

extern "C" size_t rt_cuda_multi_shard_pitch(unsigned width, unsigned height, unsigned stripRows, unsigned world) {
  unsigned maxRows = 0;
  for (unsigned g = 0; g < world; ++g) {
    const unsigned r = rt_cuda_multi_shard_rows(height, stripRows, g, world);
    if (r > maxRows) maxRows = r;
  }
  return (((size_t)maxRows * width * 3 + 15) / 16) * 16;
}

extern "C" void rt_cuda_multi_locate_row(unsigned row, unsigned stripRows, unsigned world, unsigned* rank,
                                         unsigned* localRow) {
  const unsigned strip = row / stripRows;
  if (rank) *rank = strip % world;
  if (localRow) *localRow = (strip / world) * stripRows + (row - strip * stripRows);
}

/* ---- group life cycle ---- */
static int open_local(rt_cuda_multi* m, Local& l) {
  int s = rt_cuda_init(l.device, &l.ctx);
  if (s != RT_CUDA_OK) {
    snprintf(m->lastError, sizeof m->lastError, "rt_cuda_init(device %d): %s", l.device, rt_cuda_strerror(s));
    return s;
  }
  MCU(cudaSetDevice(l.device));
  MCU(cudaEventCreate(&l.e0));
  MCU(cudaEventCreate(&l.e1));
  MCU(cudaStreamCreateWithFlags(&l.copyStream, cudaStreamNonBlocking));
  MCU(cudaEventCreateWithFlags(&l.evFrame, cudaEventDisableTiming));
  MCU(cudaEventCreateWithFlags(&l.evCopied, cudaEventDisableTiming));
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_init(int nGpus, const int* devices, rt_cuda_multi** out) {
  if (!out) return RT_CUDA_ERR_INVALID_ARG;
  *out = nullptr;
  const int have = rt_cuda_device_count();
  if (have == 0) return RT_CUDA_ERR_NO_DEVICE;
  if (nGpus < 1 || nGpus > have) return RT_CUDA_ERR_INVALID_ARG;
  rt_cuda_multi* m = new (std::nothrow) rt_cuda_multi();
  if (!m) return RT_CUDA_ERR_OUT_OF_MEMORY;
  m->world = nGpus;
  m->L.resize(nGpus);
  std::vector<int> devs(nGpus);
  for (int i = 0; i < nGpus; ++i) {
    devs[i] = devices ? devices[i] : i;
    for (int j = 0; j < i; ++j)
      if (devs[j] == devs[i]) { delete m; return RT_CUDA_ERR_INVALID_ARG; }   /* NCCL: one rank per device */
    m->L[i].device = devs[i];
    m->L[i].rank = i;
  }
  auto fail = [&](int code) { rt_cuda_multi_destroy(m); return code; };
  for (int i = 0; i < nGpus; ++i) {
    int s = open_local(m, m->L[i]);
    if (s) return fail(s);
  }
  std::vector<ncclComm_t> comms(nGpus);
  ncclResult_t r = ncclCommInitAll(comms.data(), nGpus, devs.data());
  if (r != ncclSuccess) {
    fprintf(stderr, "rt_cuda_multi_init: ncclCommInitAll: %s\n", ncclGetErrorString(r));
    return fail(RT_CUDA_ERR_NCCL);
  }
  for (int i = 0; i < nGpus; ++i) m->L[i].comm = comms[i];
  *out = m;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_unique_id(void* id, size_t idBytes) {
  if (!id || idBytes < sizeof(ncclUniqueId)) return RT_CUDA_ERR_INVALID_ARG;
  ncclUniqueId u;
  if (ncclGetUniqueId(&u) != ncclSuccess) return RT_CUDA_ERR_NCCL;
  memcpy(id, &u, sizeof u);
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_init_rank(int device, const void* id, size_t idBytes, int rank, int worldSize,
                                       rt_cuda_multi** out) {
  if (!out) return RT_CUDA_ERR_INVALID_ARG;
  *out = nullptr;
  if (!id || idBytes < sizeof(ncclUniqueId) || worldSize < 1 || rank < 0 || rank >= worldSize) return RT_CUDA_ERR_INVALID_ARG;
  if (rt_cuda_device_count() == 0) return RT_CUDA_ERR_NO_DEVICE;
  rt_cuda_multi* m = new (std::nothrow) rt_cuda_multi();
  if (!m) return RT_CUDA_ERR_OUT_OF_MEMORY;
  m->world = worldSize;
  m->L.resize(1);
  m->L[0].device = device;
  m->L[0].rank = rank;
  int s = open_local(m, m->L[0]);
  if (s) { rt_cuda_multi_destroy(m); return s; }
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclResult_t r = ncclCommInitRank(&m->L[0].comm, worldSize, u, rank);
  if (r != ncclSuccess) {
    fprintf(stderr, "rt_cuda_multi_init_rank: ncclCommInitRank: %s\n", ncclGetErrorString(r));
    rt_cuda_multi_destroy(m);
    return RT_CUDA_ERR_NCCL;
  }
  *out = m;
  return RT_CUDA_OK;
}

extern "C" void rt_cuda_multi_destroy(rt_cuda_multi* m) {
  if (!m) return;
  for (Local& l : m->L) {
    cudaSetDevice(l.device);
    if (l.ctx) rt_cuda_synchronize(l.ctx);
    if (l.comm) ncclCommDestroy(l.comm);
    cudaFree(l.dGather); cudaFree(l.dFrame);
    if (l.e0) cudaEventDestroy(l.e0);
    if (l.e1) cudaEventDestroy(l.e1);
    if (l.copyStream) { cudaStreamSynchronize(l.copyStream); cudaStreamDestroy(l.copyStream); }
    if (l.evFrame) cudaEventDestroy(l.evFrame);
    if (l.evCopied) cudaEventDestroy(l.evCopied);
    if (l.ctx) rt_cuda_destroy(l.ctx);
  }
  cudaGetLastError();
  delete m;
}

extern "C" int rt_cuda_multi_world_size(rt_cuda_multi* m) { return m ? m->world : 0; }
extern "C" int rt_cuda_multi_local_count(rt_cuda_multi* m) { return m ? (int)m->L.size() : 0; }
extern "C" rt_cuda_ctx* rt_cuda_multi_context(rt_cuda_multi* m, int local) {
  return (m && local >= 0 && local < (int)m->L.size()) ? m->L[local].ctx : nullptr;
}
extern "C" const char* rt_cuda_multi_last_error(rt_cuda_multi* m) { return m ? m->lastError : ""; }

extern "C" int rt_cuda_multi_upload_scene(rt_cuda_multi* m, const rt_sphere* spheres, unsigned sphNum,
                                          const rt_light* lights, unsigned lgtNum) {
  if (!m) return RT_CUDA_ERR_INVALID_ARG;
  for (Local& l : m->L) MRT(l, rt_cuda_upload_scene(l.ctx, spheres, sphNum, lights, lgtNum));
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_set_option(rt_cuda_multi* m, const char* key, long value) {
  if (!m) return RT_CUDA_ERR_INVALID_ARG;
  for (Local& l : m->L) MRT(l, rt_cuda_set_option(l.ctx, key, value));
  return RT_CUDA_OK;
}

template <typename T>
static int grow(rt_cuda_multi* m, Local& l, T*& p, size_t& cap, size_t need) {
  if (need > cap) {
    MRT(l, rt_cuda_synchronize(l.ctx));
    cudaFree(p); p = nullptr; cap = 0;
    MCU(cudaMalloc(&p, need));
    cap = need;
  }
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_render(rt_cuda_multi* m, unsigned width, unsigned height, float zoom,
                                    float aliasFactor, int maxStack, unsigned stripRows) {
  if (!m) return RT_CUDA_ERR_INVALID_ARG;
  if (!width || !height) return RT_CUDA_ERR_INVALID_ARG;
  const unsigned strip = stripRows ? stripRows : RT_CUDA_MULTI_DEFAULT_STRIP;
  const unsigned G = (unsigned)m->world;
  const size_t pitch = rt_cuda_multi_shard_pitch(width, height, strip, G);
  const size_t frameBytes = (size_t)width * height * 3;
  m->haveFrame = false;
  for (Local& l : m->L) {
    MCU(cudaSetDevice(l.device));
    int s = grow(m, l, l.dGather, l.gatherCap, pitch * G);
    if (s) return s;
    s = grow(m, l, l.dFrame, l.frameCap, frameBytes);
    if (s) return s;
  }
  /* 1. every rank renders its strips */
  for (Local& l : m->L) {
    MCU(cudaSetDevice(l.device));
    MCU(cudaEventRecord(l.e0, (cudaStream_t)rt_cuda_get_stream(l.ctx)));
    MRT(l, rt_cuda_render_strips(l.ctx, width, height, zoom, aliasFactor, maxStack, strip, (unsigned)l.rank, G));
  }
  /* 2. global maximum: non-negative float bits order like int32 */
  MNC(ncclGroupStart());
  for (Local& l : m->L) {
    void* dmax = rt_cuda_device_max(l.ctx);
    MNC(ncclAllReduce(dmax, dmax, 1, ncclInt32, ncclMax, l.comm, (cudaStream_t)rt_cuda_get_stream(l.ctx)));
  }
  MNC(ncclGroupEnd());
  /* 3. quantise with it, straight into this rank's block; 4. gather in place */
  for (Local& l : m->L) MRT(l, rt_cuda_quantise_to(l.ctx, l.dGather + (size_t)l.rank * pitch, pitch, 0.f));
  MNC(ncclGroupStart());
  for (Local& l : m->L)
    MNC(ncclAllGather(l.dGather + (size_t)l.rank * pitch, l.dGather, pitch, ncclUint8, l.comm,
                      (cudaStream_t)rt_cuda_get_stream(l.ctx)));
  MNC(ncclGroupEnd());
  /* 5. strips -> frame */
  for (Local& l : m->L) {
    MCU(cudaSetDevice(l.device));
    if (l.copyPending)       /* the previous frame may still be on its way to the host */
      MCU(cudaStreamWaitEvent((cudaStream_t)rt_cuda_get_stream(l.ctx), l.evCopied, 0));
    MRT(l, rt_cuda_assemble_rgb8(l.ctx, l.dGather, l.dFrame, width, height, strip, G, pitch));
    MCU(cudaEventRecord(l.e1, (cudaStream_t)rt_cuda_get_stream(l.ctx)));
  }
  m->W = width; m->H = height;
  m->haveFrame = true;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_synchronize(rt_cuda_multi* m) {
  if (!m) return RT_CUDA_ERR_INVALID_ARG;
  for (Local& l : m->L) MRT(l, rt_cuda_synchronize(l.ctx));
  return RT_CUDA_OK;
}

extern "C" void* rt_cuda_multi_device_frame(rt_cuda_multi* m, int local) {
  return (m && local >= 0 && local < (int)m->L.size()) ? (void*)m->L[local].dFrame : nullptr;
}

extern "C" int rt_cuda_multi_readback_rgb8(rt_cuda_multi* m, int local, unsigned char* dst, float* outMax) {
  if (!m || !dst || local < 0 || local >= (int)m->L.size()) return RT_CUDA_ERR_INVALID_ARG;
  if (!m->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  Local& l = m->L[local];
  MCU(cudaSetDevice(l.device));
  cudaStream_t st = (cudaStream_t)rt_cuda_get_stream(l.ctx);
  MCU(cudaMemcpyAsync(dst, l.dFrame, (size_t)m->W * m->H * 3, cudaMemcpyDeviceToHost, st));
  unsigned bits = 0;
  MCU(cudaMemcpyAsync(&bits, rt_cuda_device_max(l.ctx), sizeof bits, cudaMemcpyDeviceToHost, st));
  MCU(cudaStreamSynchronize(st));
  if (outMax) {
    float f; memcpy(&f, &bits, sizeof f);
    *outMax = (f == 0.f) ? 1.f : f;     /* algebra.h:86-88 */
  }
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_readback_rgb8_async(rt_cuda_multi* m, int local, unsigned char* dst) {
  if (!m || !dst || local < 0 || local >= (int)m->L.size()) return RT_CUDA_ERR_INVALID_ARG;
  if (!m->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  Local& l = m->L[local];
  MCU(cudaSetDevice(l.device));
  if (l.copyPending) MCU(cudaEventSynchronize(l.evCopied));
  MCU(cudaEventRecord(l.evFrame, (cudaStream_t)rt_cuda_get_stream(l.ctx)));
  MCU(cudaStreamWaitEvent(l.copyStream, l.evFrame, 0));
  MCU(cudaMemcpyAsync(dst, l.dFrame, (size_t)m->W * m->H * 3, cudaMemcpyDeviceToHost, l.copyStream));
  MCU(cudaEventRecord(l.evCopied, l.copyStream));
  l.copyPending = true;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_readback_wait(rt_cuda_multi* m, int local) {
  if (!m || local < 0 || local >= (int)m->L.size()) return RT_CUDA_ERR_INVALID_ARG;
  Local& l = m->L[local];
  if (!l.copyPending) return RT_CUDA_OK;
  MCU(cudaSetDevice(l.device));
  MCU(cudaEventSynchronize(l.evCopied));
  l.copyPending = false;
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_step_ms(rt_cuda_multi* m, int local, float* ms) {
  if (!m || !ms || local < 0 || local >= (int)m->L.size()) return RT_CUDA_ERR_INVALID_ARG;
  if (!m->haveFrame) return RT_CUDA_ERR_NO_FRAME;
  Local& l = m->L[local];
  MCU(cudaSetDevice(l.device));
  MCU(cudaEventSynchronize(l.e1));
  MCU(cudaEventElapsedTime(ms, l.e0, l.e1));
  return RT_CUDA_OK;
}

extern "C" int rt_cuda_multi_flush_l2(rt_cuda_multi* m) {
  if (!m) return RT_CUDA_ERR_INVALID_ARG;
  for (Local& l : m->L) MRT(l, rt_cuda_flush_l2(l.ctx));
  return RT_CUDA_OK;
}
