/* rt_cuda_multi.h — C-ABI of the multi-GPU render (librt_cuda_multi.so).
 *
 * The reference drives ONE device through one in-order cl_command_queue
 * (/root/reference/raytracer_gamma/main.cpp:229; launch :357-362, readback :460).
 * Pixels are independent (raytrace_kernel.cl:884-972 writes only dst[gid]) and the
 * scene is tiny, so one frame shards by rows over the GPUs of a box:
 * rank g renders the strips {k : k mod G == g} of `stripRows` rows with the scene
 * replicated.  Normalisation is by the GLOBAL maximum (algebra.h:68-91), so a frame is
 *
 *   rt_cuda_render_strips            every rank, its rows
 *   ncclAllReduce(MAX, 1 word)       the maxima are non-negative floats: their bits order as int32
 *   rt_cuda_quantise_to              main.cpp:71-76 with the global maximum, straight into this
 *                                    rank's block of the gather buffer
 *   ncclAllGather(RGB8 strips)       in place
 *   rt_cuda_assemble_rgb8            strip de-interleave -> the H x W x 3 frame on every rank
 *
 * all enqueued on each context's stream; nothing here renders, it only sequences the
 * single-GPU C-ABI (rt_cuda.h) and NCCL.  Two ways to form the group:
 *
 *   rt_cuda_multi_init       ONE process drives nGpus devices (ncclCommInitAll) — what the
 *                            reference's single-threaded main() becomes: `rt_gamma --gpus N`
 *   rt_cuda_multi_init_rank  one process per GPU (torchrun, MPI...): rank 0 makes an id with
 *                            rt_cuda_multi_unique_id, the launcher hands it to every rank
 *
 * Same conventions as rt_cuda.h: plain pointers and sizes, 0 or a negative rt_cuda_status,
 * never exits, one thread at a time per group.
 */
#ifndef RT_CUDA_MULTI_H
#define RT_CUDA_MULTI_H

#include <stddef.h>
#include <stdint.h>
#include "rt_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_cuda_multi rt_cuda_multi;

#define RT_CUDA_MULTI_ID_BYTES 128       /* sizeof(ncclUniqueId) */
#define RT_CUDA_MULTI_DEFAULT_STRIP 4u   /* rows per strip: one row of 8x4 pixel tiles */

/* Row-strip arithmetic (pure host functions, no device needed). */
unsigned rt_cuda_multi_shard_rows(unsigned height, unsigned stripRows, unsigned rank, unsigned world);
/* bytes of one rank's block in the gather buffer: the largest shard, 16-byte aligned */
size_t rt_cuda_multi_shard_pitch(unsigned width, unsigned height, unsigned stripRows, unsigned world);
/* owner of a row and its index inside the owner's packed shard */
void rt_cuda_multi_locate_row(unsigned row, unsigned stripRows, unsigned world, unsigned* rank, unsigned* localRow);

/* One process, nGpus devices.  devices == NULL means 0 .. nGpus-1. */
int rt_cuda_multi_init(int nGpus, const int* devices, rt_cuda_multi** out);

/* One process per GPU. */
int rt_cuda_multi_unique_id(void* id, size_t idBytes);
int rt_cuda_multi_init_rank(int device, const void* id, size_t idBytes, int rank, int worldSize,
                            rt_cuda_multi** out);

int rt_cuda_multi_world_size(rt_cuda_multi* m);
int rt_cuda_multi_local_count(rt_cuda_multi* m);            /* contexts driven by this process */
rt_cuda_ctx* rt_cuda_multi_context(rt_cuda_multi* m, int local);   /* for options, stats, streams */

/* Replicate the scene on every local device (rt_cuda_upload_scene each). */
int rt_cuda_multi_upload_scene(rt_cuda_multi* m, const rt_sphere* spheres, unsigned sphNum,
                               const rt_light* lights, unsigned lgtNum);
int rt_cuda_multi_set_option(rt_cuda_multi* m, const char* key, long value);

/* One frame, asynchronous on the contexts' streams.  stripRows 0 = RT_CUDA_MULTI_DEFAULT_STRIP. */
int rt_cuda_multi_render(rt_cuda_multi* m, unsigned width, unsigned height, float zoom,
                         float aliasFactor, int maxStack, unsigned stripRows);
int rt_cuda_multi_synchronize(rt_cuda_multi* m);

/* The assembled frame (H*W*3 bytes, every rank holds it) of local context `local`:
 * device pointer, or a copy to host memory (waits for the frame).  *outMax (nullable)
 * receives the global maximum, 1 when the frame is black. */
void* rt_cuda_multi_device_frame(rt_cuda_multi* m, int local);
int rt_cuda_multi_readback_rgb8(rt_cuda_multi* m, int local, unsigned char* dst, float* outMax);

/* Asynchronous copy of local context `local`'s assembled frame to dst (ideally rt_cuda_host_alloc
 * memory) on a second stream; the next rt_cuda_multi_render may be issued at once and overlaps the
 * copy (its assembly step waits for it).  One copy in flight per local context. */
int rt_cuda_multi_readback_rgb8_async(rt_cuda_multi* m, int local, unsigned char* dst);
int rt_cuda_multi_readback_wait(rt_cuda_multi* m, int local);

/* Device time of the last rt_cuda_multi_render on local context `local` (CUDA events on its
 * stream around the whole step: strips, collectives, assembly), in milliseconds. */
int rt_cuda_multi_step_ms(rt_cuda_multi* m, int local, float* ms);
/* Evict L2 on every local device (benchmark aid, see rt_cuda_flush_l2). */
int rt_cuda_multi_flush_l2(rt_cuda_multi* m);

const char* rt_cuda_multi_last_error(rt_cuda_multi* m);
void rt_cuda_multi_destroy(rt_cuda_multi* m);

#ifdef __cplusplus
}
#endif
#endif /* RT_CUDA_MULTI_H */
