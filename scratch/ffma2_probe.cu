// Throughput probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 upk(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE> __global__ void k(float *out, float s, int iters) {
  float a[16]; unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const unsigned long long sp = pk(s, s * 0.5f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(s), "f"(a[(i + 1) & 15]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(sp, p[(i + 1) & 7], p[i]);
    }
  }
  float r = 0;
  if (MODE == 0) for (int i = 0; i < 16; ++i) r += a[i]; else for (int i = 0; i < 8; ++i) { float2 v = upk(p[i]); r += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) k<0><<<148 * 8, 256>>>(d, 1.0001f, iters); else k<1><<<148 * 8, 256>>>(d, 1.0001f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = 148.0 * 8 * 256 * 16 * iters;    // scalar FMAs in both modes
    printf("%s: %.3f ms, %.1f TFMA/s (%.1f TFLOP/s), warp-instr/clk/SM = %.2f at 1.9 GHz\n", mode ? "FFMA2" : "FFMA ", ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9,
           (fmas / (mode ? 2 : 1) / 32) / (ms * 1e-3 * 1.9e9 * 148));
  }
  return 0;
}
