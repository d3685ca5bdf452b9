// Do IMAD (FMA-heavy pipe) and FFMA / FFMA2 streams overlap on sm_100a?  Times loops with I IMADs and F FP ops per iteration.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 upk(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
template <int NI, int NF, int PACKED> __global__ void k(float *out, int s, float fs, int iters) {
  int a[8]; float f[8]; unsigned long long p[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; f[i] = threadIdx.x * 0.001f + i; p[i] = pk(f[i], f[i] + 1.f); }
  const unsigned long long sp = pk(fs, fs * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u < NI) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[u]) : "r"(s), "r"(a[(u + 1) & 7]));
      if (u < NF) {
        if (PACKED) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[u]) : "l"(sp), "l"(p[(u + 1) & 7]));
        else asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[u]) : "f"(fs), "f"(f[(u + 1) & 7]));
      }
    }
  }
  float r = 0;
  for (int i = 0; i < 8; ++i) { float2 v = upk(p[i]); r += a[i] + f[i] + v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NI, int NF, int PACKED> void run(float *d, const char *name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k<NI, NF, PACKED><<<148 * 8, 256>>>(d, 3, 1.0001f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  // cycles per iteration per SMSP at 1.9 GHz: each SMSP runs 8*256/32/4 = 16 warps
  printf("%-28s %.3f ms  -> %.2f cycles per warp-iteration per SMSP (at 1.9 GHz)\n", name, best, best * 1e-3 * 1.9e9 / iters / 16);
}
int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<8, 0, 0>(d, "8 IMAD");
  run<0, 8, 0>(d, "8 FFMA");
  run<0, 8, 1>(d, "8 FFMA2");
  run<8, 8, 0>(d, "8 IMAD + 8 FFMA");
  run<8, 8, 1>(d, "8 IMAD + 8 FFMA2");
  run<8, 4, 1>(d, "8 IMAD + 4 FFMA2");
  run<4, 8, 0>(d, "4 IMAD + 8 FFMA");
  run<4, 8, 1>(d, "4 IMAD + 8 FFMA2");
  return 0;
}
