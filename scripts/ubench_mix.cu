// Development micro-benchmark: how many issue cycles do FFMA2 / FFMA / SHF / LDS mixes cost on sm_100a?
// Each kernel runs 8 independent FMA chains per iteration plus K integer funnel shifts and L broadcast LDS.128.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi){ asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned shf(unsigned a, unsigned b){ unsigned d; asm volatile("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(d) : "r"(a), "r"(b)); return d; }

#define LDS4(base, idx) do { float vx, vy, vz, vw; unsigned ad = (unsigned)__cvta_generic_to_shared(&base[idx]); \
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(vx), "=f"(vy), "=f"(vz), "=f"(vw) : "r"(ad)); \
  asm volatile("" :: "f"(vx), "f"(vy), "f"(vz), "f"(vw)); } while (0)
// PACKED: 1 = 8 FFMA2 per round (16 lane-FMAs), 0 = 16 scalar FFMA per round.  K shifts, L LDS.128 per round.
template<int PACKED, int K, int L> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, long long* cyc) {
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(a, b, a, b);
  __syncthreads();
  float r = 0.f;
  unsigned s0 = threadIdx.x, s1 = blockIdx.x, s2 = 3, s3 = 5;
  float4 acc4 = make_float4(0,0,0,0);
  long long t0 = clock64();
  if (PACKED) {
    u64 x[8]; for (int i=0;i<8;++i) x[i]=pack(threadIdx.x+i, threadIdx.x-i);
    u64 B=pack(b,b*0.999f);
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int u=0;u<8;++u) {
        u64 S = pack(a, a);
        #pragma unroll
        for (int i=0;i<8;++i) {
          x[i]=fma2(S,x[i],B);
          if (i < K) { if (i&1) s1 = shf(s0, s1); else s0 = shf(s1, s0); }
          if (i >= 8 - (K > 8 ? K - 8 : 0)) { if (i&1) s3 = shf(s2, s3); else s2 = shf(s3, s2); }
          if (i < L) { LDS4(sm, (it + u*8 + i) & 255); }
        }
      }
    }
    for (int i=0;i<8;++i){ float lo,hi; unpack(x[i],lo,hi); r+=lo+hi; }
  } else {
    float x[16]; for (int i=0;i<16;++i) x[i]=threadIdx.x+i;
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int u=0;u<8;++u) {
        #pragma unroll
        for (int i=0;i<8;++i) {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2*i]) : "f"(a), "f"(b));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2*i+1]) : "f"(a), "f"(b));
          if (i < K) { if (i&1) s1 = shf(s0, s1); else s0 = shf(s1, s0); }
          if (i >= 8 - (K > 8 ? K - 8 : 0)) { if (i&1) s3 = shf(s2, s3); else s2 = shf(s3, s2); }
          if (i < L) { LDS4(sm, (it + u*8 + i) & 255); }
        }
      }
    }
    for (int i=0;i<16;++i) r+=x[i];
  }
  long long t1 = clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=r + (float)(s0 ^ s1 ^ s2 ^ s3) + acc4.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}
template<int PACKED, int K, int L> void run(int sms, int warpsPerSched) {
  float* d; long long* c; cudaMalloc(&d, sms*8*256*sizeof(float)); cudaMalloc(&c, 8);
  const int blocks = sms * warpsPerSched / 2;       // 256 threads = 8 warps = 2 per scheduler
  const int iters = 2048;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PACKED,K,L><<<blocks,256>>>(d, iters/8, 0.999f, 0.001f, c);
  cudaEventRecord(e0); k<PACKED,K,L><<<blocks,256>>>(d, iters, 0.999f, 0.001f, c); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  // cycles per round (8 FFMA2 or 16 FFMA + K SHF + L LDS) per scheduler = cycles / (iters*8 rounds) / warpsPerSched
  double perRound = (double)hc / (iters * 8.0) / warpsPerSched;
  double tf = (double)blocks*256*(double)iters*8.0*16.0*2.0/(ms*1e-3)/1e12;
  printf("%s K=%2d L=%d warps/sched=%d : %.2f cycles per round per scheduler-warp (ideal 16), %.2f TFLOP/s\n", PACKED?"FFMA2":"FFMA ", K, L, warpsPerSched, perRound, tf);
  cudaFree(d); cudaFree(c);
}
int main(){ cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int sms=p.multiProcessorCount;
  run<1,0,0>(sms,4); run<1,2,0>(sms,4); run<1,4,0>(sms,4); run<1,8,0>(sms,4); run<1,16,0>(sms,4);
  run<0,0,0>(sms,4); run<0,2,0>(sms,4); run<0,4,0>(sms,4); run<0,8,0>(sms,4); run<0,16,0>(sms,4);
  run<1,0,1>(sms,4); run<1,0,2>(sms,4); run<1,2,1>(sms,4); run<0,0,1>(sms,4); run<0,2,1>(sms,4);
  run<1,2,1>(sms,2); run<1,2,1>(sms,6); run<1,2,1>(sms,8);
  return 0; }
