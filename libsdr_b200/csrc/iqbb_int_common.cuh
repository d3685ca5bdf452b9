// iqbb_int_common.cuh -- pieces shared by the integer IQBaseBand kernels (iqbb_kernels.cu, iqbb_warp_kernels.cu).
#pragma once
#include "iqbb_kernels.cuh"

namespace sdrg {

struct IqbbTaps { int4 t[32]; };   // Gauss-form taps {kr, ki - kr, kr + ki, 0} in the kernel's parameter (constant) bank

namespace {

__device__ __forceinline__ void unpack16(uint32_t v, int &re, int &im) {
  re = (int)(short)(v & 0xffffu);
  im = ((int)v) >> 16;
}

// One input sample as packed (re | im << 16) int16 pair.  fmt 0: complex<int16_t> as is; fmt 2/3:
// AutoCast< complex<int16_t> > fused into the load (src/autocast.hh:187-204): complex uint8 read
// through an int8_t pointer, (v - 127) << 8 (reference quirk), resp. complex int8, v << 8.
__device__ __forceinline__ uint32_t load_cs16(const void *base, int64_t idx, uint32_t fmt) {
  if (fmt == 0) return ((const uint32_t *)base)[idx];
  if (fmt == 4) return (uint32_t)((const uint16_t *)base)[idx];      // real input: imaginary part 0
  const char2 s = ((const char2 *)base)[idx];
  const int bias = fmt == 2 ? 127 : 0;
  const uint32_t re = (uint32_t)(uint16_t)(int16_t)(((int)s.x - bias) << 8);
  const uint32_t im = (uint32_t)(uint16_t)(int16_t)(((int)s.y - bias) << 8);
  return re | (im << 16);
}

// ---- shared prologue: zero the next call's accumulators, roll the history ----------------------
template <typename Sample, typename Acc>
__device__ __forceinline__ void prologue(const IqbbAccumArgs &a) {
  Acc *nxt = (Acc *)a.acc_next;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < a.zero_next; k += gridDim.x * blockDim.x) {
    Acc zero; zero.x = 0; zero.y = 0;
    nxt[k] = zero;
  }
  if (blockIdx.x == 0) {
    const int64_t H = a.hist_len, n = a.n;
    if (a.in_fmt == 5) {                                 // real int8 stream: one byte per sample
      const signed char *x = (const signed char *)a.x, *hi = (const signed char *)a.hist_in;
      signed char *ho = (signed char *)a.hist_out;
      for (int64_t k = threadIdx.x; k < H; k += blockDim.x) { const int64_t i = n - H + k; ho[k] = (i >= 0) ? x[i] : hi[H + i]; }
    } else if (sizeof(Sample) == 4 && a.in_fmt != 0) {   // fused AutoCast: the history holds raw 8-bit pairs
      const char2 *x = (const char2 *)a.x, *hi = (const char2 *)a.hist_in;
      char2 *ho = (char2 *)a.hist_out;
      for (int64_t k = threadIdx.x; k < H; k += blockDim.x) { const int64_t i = n - H + k; ho[k] = (i >= 0) ? x[i] : hi[H + i]; }
    } else {
      const Sample *x = (const Sample *)a.x;
      const Sample *hi = (const Sample *)a.hist_in;
      Sample *ho = (Sample *)a.hist_out;
      for (int64_t k = threadIdx.x; k < H; k += blockDim.x) {
        const int64_t i = n - H + k;                     // call-relative source index
        ho[k] = (i >= 0) ? x[i] : hi[H + i];
      }
    }
  }
}


__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


}  // namespace

// iqbb_warp_kernels.cu: barrier-free per-warp variant of the fixed-tap integer kernel (sub_sample >= 16, <= 32 taps).
// `kind`: 0 complex taps (Gauss form), 3 real taps (k_im == 0 for every tap), 4 real taps symmetric about the middle
// of an even count.  Returns SDRG_OK after launching, or -1 if there is no instantiation (the caller falls back).
int launch_iqbb_accum_warp(int scalar, const IqbbAccumArgs &a, const IqbbTaps &taps, int lp, cudaStream_t st);

}  // namespace sdrg
