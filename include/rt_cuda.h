/* rt_cuda.h — C-ABI of the B200 trace loop (librt_cuda.so).
 *
 * Drop-in boundary for raytracer-gamma's render call.  Each entry point replaces
 * a piece of the reference's OpenCL plumbing in main.cpp (paths relative to
 * /root/reference/raytracer_gamma/); INTEGRATION.md shows the edited main().
 *
 *   rt_cuda_init            main.cpp:182-230  platform/device pick, context, queue
 *                           (+ device_info.cpp:30-125 via rt_cuda_device_info)
 *   rt_cuda_upload_scene    main.cpp:277-294  clCreateBuffer x2 + clEnqueueWriteBuffer x2
 *   rt_cuda_render          main.cpp:339-362  clSetKernelArg 0-10 + clEnqueueNDRangeKernel + clFinish
 *                           of `__kernel raytrace` (raytrace_kernel.cl:870-973).  PARITY TARGET: the CPU
 *                           copy of the algorithm — the loop main.cpp:404-453 around rayTrace()
 *                           (raytracer.h:410) — reproduced bit for bit.  The OpenCL copy differs from it
 *                           in two places: polarisedReflection is all-float there (raytrace_kernel.cl:
 *                           399-432; double on the CPU, raytracer.h:380-393) and its RTSTACK_MAXSIZE is 5
 *                           (raytrace_kernel.cl:58; 6 on the CPU, raytraceStack.h:10).  Pass maxStack = 5
 *                           for kernel-equivalent recursion depth; the Fresnel difference stays within
 *                           the 1-LSB 8-bit tolerance.
 *   rt_cuda_readback        main.cpp:456-471  clEnqueueReadBuffer + maxColourValuePixelBuffer (algebra.h:68)
 *   rt_cuda_readback_rgb8   main.cpp:71-76    the quantiser of savePPM, on the device
 *   rt_cuda_readback_rgb8_async / rt_cuda_readback_wait   the same, overlapped with the next frame's render
 *                           (the reference blocks in clEnqueueReadBuffer(CL_TRUE), main.cpp:460)
 *   rt_cuda_destroy         main.cpp:483-489  clRelease*
 *   rt_cuda_strerror        err_code.h:31-140 err_code()
 *
 * Conventions: plain pointers and sizes only; the caller owns every host
 * pointer; the library owns all device memory until rt_cuda_destroy; every call
 * returns RT_CUDA_OK (0) or a negative rt_cuda_status and never exits the
 * process (the reference's checkError prints and exits, err_code.h:142-155 —
 * that stays the host program's choice).  Calls on one context must come from
 * one thread at a time.  There is no CPU fallback: without a CUDA device
 * rt_cuda_init fails with RT_CUDA_ERR_NO_DEVICE.
 */
#ifndef RT_CUDA_H
#define RT_CUDA_H

#include <stddef.h>
#include <stdint.h>
#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_cuda_ctx rt_cuda_ctx;

typedef enum rt_cuda_status {
  RT_CUDA_OK = 0,
  RT_CUDA_ERR_INVALID_ARG = -1,
  RT_CUDA_ERR_NO_DEVICE = -2,
  RT_CUDA_ERR_CUDA = -3,          /* a CUDA runtime call failed; see rt_cuda_last_error */
  RT_CUDA_ERR_NO_SCENE = -4,
  RT_CUDA_ERR_NO_FRAME = -5,
  RT_CUDA_ERR_TOO_LARGE = -6,
  RT_CUDA_ERR_OUT_OF_MEMORY = -7,
  RT_CUDA_ERR_NCCL = -8           /* multi-GPU library only (rt_cuda_multi.h) */
} rt_cuda_status;

/* Limits of this implementation */
#define RT_CUDA_MAX_SPHERES 12288u   /* filter records must fit one SM's shared memory */
#define RT_CUDA_MAX_STACK   16       /* RTSTACK_MAXSIZE values accepted by rt_cuda_render */
#define RT_CUDA_MAX_LIGHTS  32767u   /* the slot record keeps the light index in 15 bits */

/* Counters and timings of the last render (device-side tallies). */
typedef struct rt_cuda_stats {
  uint64_t rays;            /* calcIntersection calls of the reference algorithm (raytracer.h:145) */
  uint64_t shadow_rays;     /* of which shadow rays (raytracer.h:272)                               */
  uint64_t contain_queries; /* primaryContainer calls (raytracer.h:245)                             */
  uint64_t contain_tests;   /* its loop iterations, counted as the reference executes them          */
  uint64_t exact_tests;     /* candidates that went through the exact expressions                   */
  uint64_t samples;         /* rayTrace calls (main.cpp:439)                                        */
  uint64_t lane_iters;      /* sub-query capacity of all sphere passes (lanes x sub-queries per lane)      */
  uint64_t active_lane_iters; /* sub-queries actually served                                       */
  uint64_t served_trace, served_shadow, served_contain;   /* ... by kind                           */
  uint64_t passes;          /* warp-level passes over the sphere list                               */
  uint64_t passes_trace, passes_shadow2, passes_shadow4, passes_contain;   /* ... by loop variant   */
  uint64_t phase_cycles[6]; /* -DRT_PHASE_TIMING builds only: warp cycles in refill+vote, set-up, filter loop, resolve, advance */
  uint64_t filter_tests;    /* discriminant filter tests executed = lane_iters * padded sphere count */
  uint64_t null_rays;       /* rays with direction 0 (total internal reflection): certain miss, no sphere loop */
  uint32_t sph_num, sph_padded, lgt_num;
  uint32_t width, height, local_rows;
  float    kernel_ms;       /* trace kernel only, CUDA events on the context's stream               */
  float    max_colour;      /* NaN-skipping max of this context's rows, 0 if all black              */
  uint32_t kernel_launches; /* kernels this library launched since the last render began            */
  uint32_t grid, block, smem_bytes, staging;   /* launch shape; staging 1 = constant bank (launch parameter), 2 = shared via TMA bulk */
  uint32_t engine;          /* 1 = persistent multi-slot kernel (the only engine) */
  uint32_t accel;           /* 1 = this frame used the two-level cluster filter (option "accel")    */
  uint32_t clusters;        /* ... over this many sphere clusters                                   */
  uint32_t slots_on_chip;   /* slot records per lane held in shared memory (4 or 3), 0 = they lived in local memory */
} rt_cuda_stats;

/* Open device `device` (cudaSetDevice ordinal).  *out receives the context. */
int rt_cuda_init(int device, rt_cuda_ctx** out);

/* Copy the scene (reference AoS layout) to the device and build the kernel's SoA
 * form.  sphNum may be 0 (everything misses); lgtNum may be 0. */
int rt_cuda_upload_scene(rt_cuda_ctx* ctx, const rt_sphere* spheres, unsigned sphNum,
                         const rt_light* lights, unsigned lgtNum);

/* Render the whole width x height frame.  zoom = kZoom, aliasFactor =
 * kAliasFactor (raytrace_kernel.cl:876-877); maxStack = RTSTACK_MAXSIZE
 * (raytraceStack.h:10; 6 in the reference CPU build, 5 in the OpenCL kernel).
 * Asynchronous on the context's stream. */
int rt_cuda_render(rt_cuda_ctx* ctx, unsigned width, unsigned height, float zoom,
                   float aliasFactor, int maxStack);

/* Render only the rows r with (r / stripRows) % stripStride == stripFirst — the
 * row-strip shard of one GPU.  The context's framebuffer then holds those rows
 * packed in increasing r. */
int rt_cuda_render_strips(rt_cuda_ctx* ctx, unsigned width, unsigned height, float zoom,
                          float aliasFactor, int maxStack, unsigned stripRows,
                          unsigned stripFirst, unsigned stripStride);

/* Wait for the render, copy the float RGB rows (12 B per pixel, the
 * reference's `Vec dst[]`) into dst, and return the NaN-skipping maximum
 * (algebra.h:68-91; 1 when all black) in *outMax if non-NULL. */
int rt_cuda_readback(rt_cuda_ctx* ctx, rt_vec* dst, float* outMax);

/* Quantise on the device exactly as savePPM does (main.cpp:71-76) and copy 3
 * bytes per pixel into dst.  maxColour <= 0 uses this context's own maximum. */
int rt_cuda_readback_rgb8(rt_cuda_ctx* ctx, unsigned char* dst, float maxColour);

/* Device-side quantise only (result stays in rt_cuda_device_rgb8). */
int rt_cuda_quantise(rt_cuda_ctx* ctx, float maxColour);
/* ... into caller-provided device memory (4-byte aligned, >= localRows*W*3 bytes), e.g. this
 * rank's block of an all-gather buffer. */
int rt_cuda_quantise_to(rt_cuda_ctx* ctx, void* devDst, size_t dstBytes, float maxColour);

/* Asynchronous readback.  Quantises the frame just rendered on the
 * context's stream and copies it to dst on a second stream; returns at once with a ticket
 * (0 or 1).  dst is complete after rt_cuda_readback_wait(ctx, ticket).  The next
 * rt_cuda_render may be issued immediately: it overlaps the copy.  Two readbacks can be
 * in flight; a third call first waits for the oldest.  dst from rt_cuda_host_alloc
 * (pinned) is written by DMA directly; any other memory is staged and filled by the wait. */
int rt_cuda_readback_rgb8_async(rt_cuda_ctx* ctx, unsigned char* dst, float maxColour, int* ticket);
int rt_cuda_readback_wait(rt_cuda_ctx* ctx, int ticket);
void* rt_cuda_host_alloc(size_t bytes);          /* page-locked host memory, NULL on failure */
void  rt_cuda_host_free(void* p);

/* Device pointers for multi-GPU plumbing (NCCL gather / max all-reduce). */
void* rt_cuda_device_packed(rt_cuda_ctx* ctx);   /* float[localRows*W*3], valid after rt_cuda_pack   */
void* rt_cuda_device_rgb8(rt_cuda_ctx* ctx);     /* uint8[localRows*W*3], valid after rt_cuda_quantise */
void* rt_cuda_device_max(rt_cuda_ctx* ctx);      /* one float (bits), 0 if black                     */
int   rt_cuda_pack(rt_cuda_ctx* ctx);            /* float4 framebuffer -> packed float RGB on device */

/* Use an existing cudaStream_t (e.g. the framework's current stream). */
int rt_cuda_set_stream(rt_cuda_ctx* ctx, void* cudaStream);
void* rt_cuda_get_stream(rt_cuda_ctx* ctx);      /* the cudaStream_t the context launches on */
int rt_cuda_get_device(rt_cuda_ctx* ctx);
int rt_cuda_synchronize(rt_cuda_ctx* ctx);
/* Benchmark aid: overwrite a 256 MiB scratch buffer on the context's stream (evicts the 126 MB L2). */
int rt_cuda_flush_l2(rt_cuda_ctx* ctx);

/* Tuning / debug switches: "staging" 0 auto (= 2) | 1 constant bank: the filter records travel as a
 *   __grid_constant__ launch parameter, <= 1024 spheres | 2 shared memory filled by TMA bulk copies;
 * "no_filter" 1 = exact test against every sphere; "blocks_per_sm" 0 auto;
 * "order" 0 auto (= 1 on frames large enough for it to pay, else 2) | 1: every tile's first group, then the tiles it showed to be deep, then the rest (so that a
 *   launch does not end on a few warps working off 100-query chains) | 2: tiles in scanline order;
 *   "deep_at" / "sweep_step": its tuning (0 = default: 24 queries / at most one tile per claim);
 * "lockstep" 0 auto (by the number of filter records: the whole CTA up to 384, halves of it up to 896, off beyond) |
 *   1 whole CTA | 2 off | 3 half-CTAs: the warps of a group vote and start their passes together, which keeps the
 *   instruction working set of an SM inside its instruction cache;
 * "slot_mode" 0 auto (slot records in shared memory whenever two CTAs per SM still fit) | 1 shared | 2 local;
 * (development builds, -DRT_DEV_VARIANTS: "slots" 3 | 4 for the local-memory kernel);
 * "accel" 0 off | 1 = two-level cluster filter where it pays (>= 768 spheres) | 2 = from 32 spheres.
 *   Same frame bit for bit; fewer filter tests.  Off by default: the reference's
 *   algorithm is brute force and the default kernel is measured against that roofline. */
int rt_cuda_set_option(rt_cuda_ctx* ctx, const char* key, long value);

int rt_cuda_get_stats(rt_cuda_ctx* ctx, rt_cuda_stats* out);

/* Multi-GPU assembly on the context's stream: `gathered` (device) holds nShards blocks of
 * shardPitchBytes, block g = the RGB8 rows of shard g (rt_cuda_render_strips with
 * stripFirst = g, stripStride = nShards) packed in increasing row order, e.g. the output
 * of an NCCL all-gather.  Writes the height x width x 3 frame to `out` (device). */
int rt_cuda_assemble_rgb8(rt_cuda_ctx* ctx, const void* gathered, void* out, unsigned width,
                          unsigned height, unsigned stripRows, unsigned nShards,
                          size_t shardPitchBytes);

/* Measure the FP32 FMA ceiling of this device (register-only FFMA chains), in TFLOP/s:
 * the empirical denominator beside the nominal #SM x 128 x 2 x clock. */
int rt_cuda_ffma_peak(rt_cuda_ctx* ctx, int iters, float* outTflops);

void rt_cuda_destroy(rt_cuda_ctx* ctx);

const char* rt_cuda_strerror(int status);
const char* rt_cuda_last_error(rt_cuda_ctx* ctx);   /* text of the last CUDA failure */

int rt_cuda_device_count(void);
/* One-paragraph description of a device (the analogue of output_device_info). */
int rt_cuda_device_info(int device, char* buf, size_t bufSize);

#ifdef __cplusplus
}
#endif
#endif /* RT_CUDA_H */
