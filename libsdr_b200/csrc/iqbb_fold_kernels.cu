// iqbb_fold_kernels.cu -- IQBaseBand<float>: the folded (HBM-bound) accumulate kernel.
//
// The float path is linear (no truncating shifts), so FIR -> NCO -> boxcar can be folded into one
// weight per input sample.  With c(n) = lut[idx(n)] the NCO factor and
// window W_m = [.., n_hi], sample p contributes  x[p] * sum_{n in W_m, p <= n <= p+L-1} c(n) k[L-1-(n-p)].
// The 15-bit phase factorises exactly: phi_p = 256 a + r  =>  c(p+d) = A(a) * B(r,d) with
// A(a) = exp(-/+ 2 pi i a/128) and B depending only on (r, d); hence
//     full weight      G(p)   = A(a_p) * U(r_p, 0)
//     tail past a window end at distance e = n_hi+1-p (1 <= e <= L-1):
//                      T_e(p) = A(a_b) * U(r_b, e),   phi_b = phase at the next window's first sample
//     U(r,e) = sum_{j=0}^{L-1-e} B(r,j) k[L-1-e-j]       (256 x L table, built on the host in double)
// Each sample is read ONCE and contributes (G - T) x to its own window and T x to the next one
// (requires ss >= L-1); no halo re-reads, no FIR history.  ~10 flop and 8 bytes per input sample.
//
// Mapping: a warp owns a contiguous segment of the call's samples and walks the windows that cross
// it; lanes stride over the samples (coalesced 8-byte loads, 4 independent loads in flight per
// lane), keep per-lane partial sums, and only at a window end reduce with shuffles and issue one
// RED.ADD per component into the per-call window accumulators (same finalize kernel as the direct
// path).  The "next window" partials simply become the running partials of the next window.
#include "iqbb_kernels.cuh"

namespace sdrg {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kFoldThreads = 256;
constexpr int kFoldWarps = kFoldThreads / 32;

__device__ __forceinline__ float2 ld_stream(const float2 *p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(float2 &acc, float2 w, float2 x) {
  acc.x = fmaf(w.x, x.x, acc.x); acc.x = fmaf(-w.y, x.y, acc.x);
  acc.y = fmaf(w.x, x.y, acc.y); acc.y = fmaf(w.y, x.x, acc.y);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}
__device__ __forceinline__ void flush(float *acc, uint32_t slot, float2 v, int lane) {
  const float sr = warp_sum(v.x), si = warp_sum(v.y);
  if (lane == 0 && (sr != 0.f || si != 0.f)) { atomicAdd(acc + 2 * (size_t)slot, sr); atomicAdd(acc + 2 * (size_t)slot + 1, si); }
}

__global__ void __launch_bounds__(kFoldThreads) iqbb_fold_f32_kernel(const IqbbFoldArgs a) {
  __shared__ float2 sA[128];
  __shared__ float2 sH[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // zero the next call's accumulators (see iqbb_kernels.cu)
  for (uint32_t k = blockIdx.x * blockDim.x + tid; k < a.zero_next; k += gridDim.x * blockDim.x)
    ((float2 *)a.acc_next)[k] = make_float2(0.f, 0.f);
  if (tid < 128) sA[tid] = a.tab_a[tid];
  sH[tid] = a.tab_u[(size_t)tid * a.taps_len];              // U(r, 0)
  __syncthreads();

  const uint64_t seg_lo64 = ((uint64_t)blockIdx.x * kFoldWarps + warp) * a.seg;
  if (seg_lo64 >= a.n) return;
  const uint32_t seg_lo = (uint32_t)seg_lo64;
  const uint32_t seg_hi = (uint32_t)min((uint64_t)a.n, seg_lo64 + a.seg);
  const float2 *__restrict__ x = (const float2 *)a.x;
  float *acc_out = (float *)a.acc_cur;
  const int64_t L1 = (int64_t)a.taps_len - 1;
  const int64_t win_off = (int64_t)a.first - (int64_t)a.r0;  // end(s) = (s+1)*ss + win_off

  uint32_t s = (uint32_t)(((uint64_t)a.r0 + seg_lo - ((a.first && seg_lo > 0) ? 1u : 0u)) / a.ss);
  float2 acc = make_float2(0.f, 0.f), nxt = make_float2(0.f, 0.f);
  uint32_t i = seg_lo;
  while (i < seg_hi) {
    const int64_t full_end = (int64_t)((uint64_t)(s + 1) * a.ss) + win_off;   // exclusive, call-relative
    const uint32_t wend = (uint32_t)min(full_end, (int64_t)seg_hi);
    const int64_t tr_lo = full_end - L1;                                       // first sample with a tail
    const uint32_t int_hi = (uint32_t)max((int64_t)i, min((int64_t)wend, tr_lo));

    // interior samples: weight G(p) = A(a_p) U(r_p,0)
    uint32_t j = i + lane;
    for (; j + 96 < int_hi; j += 128) {
      const float2 x0 = ld_stream(x + j), x1 = ld_stream(x + j + 32), x2 = ld_stream(x + j + 64), x3 = ld_stream(x + j + 96);
      const uint32_t p0 = (a.phase0 + j * a.inc) & 0x7fffu, p1 = (p0 + 32 * a.inc) & 0x7fffu,
                     p2 = (p0 + 64 * a.inc) & 0x7fffu, p3 = (p0 + 96 * a.inc) & 0x7fffu;
      cfma(acc, cmul(sA[p0 >> 8], sH[p0 & 255]), x0);
      cfma(acc, cmul(sA[p1 >> 8], sH[p1 & 255]), x1);
      cfma(acc, cmul(sA[p2 >> 8], sH[p2 & 255]), x2);
      cfma(acc, cmul(sA[p3 >> 8], sH[p3 & 255]), x3);
    }
    for (; j < int_hi; j += 32) {
      const float2 x0 = ld_stream(x + j);
      const uint32_t p0 = (a.phase0 + j * a.inc) & 0x7fffu;
      cfma(acc, cmul(sA[p0 >> 8], sH[p0 & 255]), x0);
    }
    // trailing samples of the window: split between this window and the next
    if ((int64_t)wend > tr_lo) {
      const uint32_t pb = (a.phase0 + (uint32_t)full_end * a.inc) & 0x7fffu;
      const float2 Ab = sA[pb >> 8];
      const float2 *__restrict__ urow = a.tab_u + (size_t)(pb & 255) * a.taps_len;
      for (j = int_hi + lane; j < wend; j += 32) {
        const float2 x0 = ld_stream(x + j);
        const uint32_t p0 = (a.phase0 + j * a.inc) & 0x7fffu;
        const float2 g = cmul(sA[p0 >> 8], sH[p0 & 255]);
        const float2 t = cmul(Ab, __ldg(urow + (uint32_t)(full_end - (int64_t)j)));
        cfma(acc, make_float2(g.x - t.x, g.y - t.y), x0);
        cfma(nxt, t, x0);
      }
    }
    i = wend;
    if ((int64_t)wend == full_end) {       // window complete within this segment
      flush(acc_out, s, acc, lane);
      acc = nxt; nxt = make_float2(0.f, 0.f);
      ++s;
    }
  }
  flush(acc_out, s, acc, lane);
  flush(acc_out, s + 1, nxt, lane);
}

}  // namespace

int launch_iqbb_fold(const IqbbFoldArgs &a, cudaStream_t st) {
  if (a.n == 0) return SDRG_OK;
  const uint64_t per_block = (uint64_t)a.seg * kFoldWarps;
  const unsigned grid = (unsigned)((a.n + per_block - 1) / per_block);
  iqbb_fold_f32_kernel<<<grid, kFoldThreads, 0, st>>>(a);
  SDRG_CHECK_LAUNCH("iqbb_fold_f32_kernel");
  return SDRG_OK;
}

}  // namespace sdrg
