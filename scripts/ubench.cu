// Development micro-benchmarks: FP32 FMA issue rates on sm_100a (scalar FFMA vs packed FFMA2).
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi){ asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template<int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float r = 0.f;
  if (MODE == 0) {          // scalar FFMA, 8 chains
    float x[8]; for (int i=0;i<8;++i) x[i]=threadIdx.x+i;
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int u=0;u<16;++u) { 
        #pragma unroll
        for (int i=0;i<8;++i) x[i]=__fmaf_rn(x[i],a,b); }
    }
    for (int i=0;i<8;++i) r+=x[i];
  } else if (MODE == 1) {   // FFMA2, 8 packed chains (16 FMAs per round), operands: pair * pair + pair
    u64 x[8]; for (int i=0;i<8;++i) x[i]=pack(threadIdx.x+i, threadIdx.x-i);
    u64 A=pack(a,a*1.0001f), B=pack(b,b*0.999f);
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int u=0;u<16;++u) {
        #pragma unroll
        for (int i=0;i<8;++i) x[i]=fma2(x[i],A,B); }
    }
    for (int i=0;i<8;++i){ float lo,hi; unpack(x[i],lo,hi); r+=lo+hi; }
  } else if (MODE == 2) {   // FFMA2 with scalar-broadcast first operand (the sphere-component form)
    u64 x[8]; for (int i=0;i<8;++i) x[i]=pack(threadIdx.x+i, threadIdx.x-i);
    u64 B=pack(b,b*0.999f);
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int u=0;u<16;++u) {
        u64 S = pack(a, a);
        #pragma unroll
        for (int i=0;i<8;++i) x[i]=fma2(S,x[i],B); }
    }
    for (int i=0;i<8;++i){ float lo,hi; unpack(x[i],lo,hi); r+=lo+hi; }
  } else if (MODE == 3) {   // 3 distinct register operands per scalar FFMA (no reuse): RF-bandwidth probe
    float x[8], y[8], z[8]; for (int i=0;i<8;++i){ x[i]=threadIdx.x+i; y[i]=a+i; z[i]=b-i; }
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int u=0;u<16;++u) {
        #pragma unroll
        for (int i=0;i<8;++i) x[i]=__fmaf_rn(y[i],z[(i+3)&7],x[i]); }
    }
    for (int i=0;i<8;++i) r+=x[i];
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int MODE> double run(int sms, int iters, double fmasPerRound) {
  float* d; cudaMalloc(&d, sms*8*256*sizeof(float));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms*8,256>>>(d, iters/8, 0.999f, 0.001f);
  cudaEventRecord(e0); k<MODE><<<sms*8,256>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); cudaFree(d);
  return (double)sms*8*256*(double)iters*16.0*fmasPerRound*2.0/(ms*1e-3)/1e12;
}
int main(){ cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int sms=p.multiProcessorCount;
  printf("scalar FFMA (reuse operands)      : %.2f TFLOP/s\n", run<0>(sms,8192,8));
  printf("FFMA2 pair*pair+pair               : %.2f TFLOP/s\n", run<1>(sms,8192,16));
  printf("FFMA2 scalar-broadcast*pair+pair   : %.2f TFLOP/s\n", run<2>(sms,8192,16));
  printf("scalar FFMA 3 distinct reg operands: %.2f TFLOP/s\n", run<3>(sms,8192,8));
  return 0; }
