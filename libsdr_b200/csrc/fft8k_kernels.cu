// fft8k_kernels.cu -- the 8192 / 4096-point transforms and the block-4096 FFT convolution (BASELINE config 3),
// restructured around three facts measured on the radix-16 Stockham kernels they replace
// (profiles/r01_final_fft_filter_kernel_summary.md: 64-register cap with spills, address/twiddle arithmetic on
// the ALU pipe costing more issue slots than the butterflies, 3 + 3 shared-memory round trips):
//
//  * An 8192-point DFT is split by ONE decimation-in-frequency radix-2 step, applied while the samples are
//    loaded, into two independent 4096-point DFTs (even and odd bins); 4096 = 16 x 16 x 16, so each half is
//    three radix-16 stages, one butterfly per thread and stage (256 threads, 16 + 16 elements per thread,
//    128 registers at 2 CTAs/SM: no spills).
//  * The stages are IN PLACE in the Cooley-Tukey sense: a butterfly reads and writes the same 16 locations, so
//    a stage needs one CTA barrier and a thread may run its two butterflies (one per half) back to back.  The
//    spectrum therefore comes out digit-reversed -- which a convolution never has to undo: the filter's
//    spectrum is stored in the same permuted order (host, config time) and the inverse transform is the mirror
//    image (decimation in time) that consumes digit-reversed input and delivers natural order.  The last
//    forward stage, the spectrum multiply and the first inverse stage all happen in registers: 4 shared-memory
//    round trips per block instead of 6.  For the plain FFTPlan transform the digit reversal is absorbed by
//    the last stage's thread mapping: thread t holds bins t + 256 q of both halves, i.e. X[2(t + 256 q)] and
//    X[2(t + 256 q) + 1], and stores them as one coalesced float4.
//  * Every twiddle is a table entry (double precision on the host, rounded once): w_4096^(q u), w_256^(q c),
//    w_8192^u, w_32^a live in shared memory, indexed so that a warp reads consecutive or broadcast words.  No
//    product chains, no sincos, no per-stage index arithmetic beyond shifts and adds.
//
// Index algebra (checked against numpy.fft by scratch/fft8k_model.py, which mirrors this file line by line):
//   half h, input index j = 256 a + 16 b + c, output bin m = qa + 16 qb + 256 qc,
//   w_4096^(j m) = w_16^(a qa) . w_4096^(qa (16 b + c)) . w_16^(b qb) . w_256^(c qb) . w_16^(c qc)
//   stage 1: DFT16 over a, twiddle w_4096^(qa (16b + c));  stage 2: DFT16 over b, twiddle w_256^(qb c);
//   stage 3: DFT16 over c.  Position 256 qa + 16 qb + qc then holds U[qa + 16 qb + 256 qc].
// Reference interfaces replaced: FFTPlan<float> (src/fftplan_fftw3.hh:79-142, FFTW3 underneath) and
// FilterSink/FilterSource (src/filternode.hh:81-88,164-181).
#include "fft_kernels.cuh"
#include "fft_device.cuh"

#include <atomic>
#include <cmath>
#include <vector>

namespace sdrg {
namespace {

constexpr int kT = 256;                       // threads per CTA
constexpr int kHalf = 4096;
// shared-memory position of element i of a half: two pads per 16 and two more per 256, so that lanes striding
// by 1 (stages 1, 2: 64-bit accesses) or by 256 (stage 3: 128-bit accesses of 16 consecutive elements, whose
// first position 290 qa + 18 qb is even, i.e. 16-byte aligned) all fall on distinct banks
__device__ __forceinline__ int pos(int i) { return i + 2 * (i >> 4) + 2 * (i >> 8); }
// the same written per digit, i = 256 A + 16 B + C: every access below is then [thread base + immediate]
__host__ __device__ constexpr int pos3(int A, int B, int C) { return 290 * A + 18 * B + C; }
constexpr int kHalfPad = kHalf + 2 * (kHalf / 16) + 2 * (kHalf / 256) + 8;

// table layout (float2 entries, forward sign exp(-2 pi i k / n))
constexpr int kT1 = 0;                        // [q][u]  w_4096^(q u), q < 16, u < 256
constexpr int kT2 = kT1 + 16 * 256;           // [q][c]  w_256^(q c),  q < 16, c < 16
constexpr int kT8 = kT2 + 16 * 16;            // [u]     w_8192^u,     u < 256
constexpr int kTabLen = kT8 + 256;
constexpr size_t kSmemBytes = (size_t)(2 * kHalfPad + kTabLen) * sizeof(float2);

__device__ __forceinline__ void load_tables(float2 *tab, const float2 *__restrict__ g) {
  for (int i = threadIdx.x; i < kTabLen; i += kT) tab[i] = g[i];
}

// w_8192^(256 a + t) = w_32^a . w_8192^t; the 16 values of w_32^a are immediates
__device__ __forceinline__ float2 root8k(const float2 wt, const int a) {
  constexpr float c32[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f,
                             3.826834324e-01f, 1.950903220e-01f, 0.0f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f,
                             -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
  constexpr float s32[16] = {0.000000000e+00f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f, 7.071067812e-01f, 8.314696123e-01f,
                             9.238795325e-01f, 9.807852804e-01f, 1.0f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f,
                             7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f};
  if (a == 0) return wt;
  if (a == 8) return make_float2(wt.y, -wt.x);
  // wt * (c - i s)
  return make_float2(wt.x * c32[a] + wt.y * s32[a], wt.y * c32[a] - wt.x * s32[a]);
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- the decimation-in-frequency pass (forward; INV conjugates every root => unnormalised backward DFT) ----
// Both halves go through a stage together: twice the independent work per thread, and every table twiddle is
// loaded once for the two butterflies that need it.
// stage 1: e[a] = u0[256 a + t], o[a] = u1[256 a + t] on entry (thread t = 16 b + c)
// The CTA barrier that frees the buffers (everybody has finished READING the previous item's last stage) sits between
// the butterflies and the stores, so the global-load latency and the first DFTs of an item overlap the tail of the
// previous one.
template <bool INV>
__device__ __forceinline__ void dif_stage1(float2 *e, float2 *o, float2 *H0, const float2 *tab, const int t) {
  const int pt = pos3(0, t >> 4, t & 15);
  dft16<INV>(e);
  dft16<INV>(o);
  __syncthreads();
  H0[pt] = e[0];
  H0[pt + kHalfPad] = o[0];
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    const float2 w = tab[kT1 + 256 * q + t];
    H0[pt + pos3(q, 0, 0)] = cmulw<INV>(e[q], w);
    H0[pt + pos3(q, 0, 0) + kHalfPad] = cmulw<INV>(o[q], w);
  }
}
// stage 2: thread t = 16 qa + c works on positions 256 qa + 16 b + c of both halves
template <bool INV>
__device__ __forceinline__ void dif_stage2(float2 *H0, const float2 *tab, const int t) {
  const int c = t & 15;
  float2 *B = H0 + pos3(t >> 4, 0, c);
  float2 e[16], o[16];
#pragma unroll
  for (int b = 0; b < 16; ++b) { e[b] = B[pos3(0, b, 0)]; o[b] = B[pos3(0, b, 0) + kHalfPad]; }
  dft16<INV>(e);
  dft16<INV>(o);
  B[0] = e[0];
  B[kHalfPad] = o[0];
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    const float2 w = tab[kT2 + 16 * q + c];
    B[pos3(0, q, 0)] = cmulw<INV>(e[q], w);
    B[pos3(0, q, 0) + kHalfPad] = cmulw<INV>(o[q], w);
  }
}
// stage 3: thread t = qa + 16 qb reads the 16 consecutive positions of 256 qa + 16 qb + c (128-bit loads);
// on return v[qc] = U[t + 256 qc]
__device__ __forceinline__ int stage3_pos(const int t) { return pos3(t & 15, t >> 4, 0); }
// The convolution does not care in which order the bins come out, so its stage 3 uses thread t = 16 qa + qb instead:
// the 16 positions it reads were all written (stage 2, threads 16 qa + c) by lanes of the SAME warp, and the inverse
// mirror holds too -- two of the five CTA barriers per block become __syncwarp().  Bins held: qa + 16 qb + 256 qc.
__device__ __forceinline__ int stage3_pos_conv(const int t) { return pos3(t >> 4, t & 15, 0); }
__device__ __forceinline__ void load16(float2 *v, const float2 *H, const int p0) {
  const float4 *h4 = (const float4 *)(H + p0);
#pragma unroll
  for (int k = 0; k < 8; ++k) { const float4 q = h4[k]; v[2 * k] = make_float2(q.x, q.y); v[2 * k + 1] = make_float2(q.z, q.w); }
}
__device__ __forceinline__ void store16(const float2 *v, float2 *H, const int p0) {
  float4 *h4 = (float4 *)(H + p0);
#pragma unroll
  for (int k = 0; k < 8; ++k) h4[k] = make_float4(v[2 * k].x, v[2 * k].y, v[2 * k + 1].x, v[2 * k + 1].y);
}

// ---- FFTPlan<float>: batched 8192-point (SPLIT) or 4096-point (two transforms per iteration) DFT ----------
template <bool INV, bool SPLIT>
__global__ void __launch_bounds__(kT, 2) fft8k_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, const int batch,
                                                      const float2 *__restrict__ gtab) {
  extern __shared__ __align__(16) unsigned char fft8k_smem[];
  float2 *H0 = (float2 *)fft8k_smem, *H1 = H0 + kHalfPad, *tab = H1 + kHalfPad;
  const int t = threadIdx.x;
  load_tables(tab, gtab);
  __syncthreads();
  const int items = SPLIT ? batch : (batch + 1) / 2;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const size_t f0 = SPLIT ? (size_t)it : (size_t)2 * it;
    const bool second = SPLIT || (f0 + 1 < (size_t)batch);
    const float2 *xa = in + f0 * (SPLIT ? 8192 : 4096);
    const float2 *xb = SPLIT ? xa + kHalf : (second ? xa + kHalf : xa);
    {
      float2 e[16], o[16];
#pragma unroll
      for (int a = 0; a < 16; ++a) { e[a] = xa[256 * a + t]; o[a] = xb[256 * a + t]; }
      if (it + (int)gridDim.x < items) {      // the next item of this CTA: 64 KB, two 128-byte lines per thread, into L2
        const char *nx = (const char *)(xa + (size_t)gridDim.x * 8192);
        prefetch_l2(nx + 128 * t); prefetch_l2(nx + 128 * (t + 256));
      }
      if (SPLIT) {
        const float2 wt = tab[kT8 + t];
#pragma unroll
        for (int a = 0; a < 16; ++a) {
          const float2 p = e[a], c = o[a];
          e[a] = caddf(p, c);
          o[a] = cmulw<INV>(csubf(p, c), root8k(wt, a));
        }
      }
      dif_stage1<INV>(e, o, H0, tab, t);
    }
    __syncthreads();
    dif_stage2<INV>(H0, tab, t);
    __syncthreads();
    {
      float2 e[16], o[16];
      const int p0 = stage3_pos(t);
      load16(e, H0, p0); load16(o, H1, p0);
      dft16<INV>(e); dft16<INV>(o);
      if (SPLIT) {               // X[2 m] = E[m], X[2 m + 1] = O[m], m = t + 256 q: one float4 per bin pair
        float4 *y = (float4 *)(out + f0 * 8192);
#pragma unroll
        for (int q = 0; q < 16; ++q) y[t + 256 * q] = make_float4(e[q].x, e[q].y, o[q].x, o[q].y);
      } else {
        float2 *ya = out + f0 * 4096, *yb = ya + kHalf;
#pragma unroll
        for (int q = 0; q < 16; ++q) { ya[t + 256 * q] = e[q]; if (second) yb[t + 256 * q] = o[q]; }
      }
    }                            // (stage 3 only read; the barrier inside the next item's stage 1 orders its stores behind this)
  }
}

// ---- block-4096 overlap-save convolution, fused ------------------------------------------------------------
// y[N + j] = IDFT_8192(DFT_8192([prev | cur]) K)[N + j] / 8192 = v0[j] - conj(w_8192^j) v1[j], v_h the 4096-point
// backward DFTs of the even / odd bins.  K arrives permuted and pre-scaled: kp[(h 16 + qc) 256 + t] = K[2 m + h] / 8192,
// m = (t >> 4) + 16 (t & 15) + 256 qc the bin thread t holds after stage 3.
// BANK = false: one filter; the last forward stage, the spectrum multiply and the first inverse stage stay in registers.
// BANK = true: F filters on one FilterSink (src/filternode.hh:262-270).  The forward transform runs once per block; its
// spectrum (digit-reversed, 64 KB) goes to a CTA-private scratch line in global memory -- written and read back by the
// SAME thread, so no barrier is involved and, with at most 2 x 148 CTAs, the scratch (19 MB) never leaves L2 -- and
// every filter then runs the inverse half.
template <bool BANK>
__global__ void __launch_bounds__(kT, 2) conv8k_kernel(const FilterArgs a, const int n_blocks) {
  extern __shared__ __align__(16) unsigned char fft8k_smem[];
  float2 *H0 = (float2 *)fft8k_smem, *H1 = H0 + kHalfPad, *tab = H1 + kHalfPad;
  const int t = threadIdx.x;
  load_tables(tab, (const float2 *)a.tab8k);
  __syncthreads();
  const float2 *x = (const float2 *)a.x;
  float2 *scratch = BANK ? (float2 *)a.spec + (size_t)blockIdx.x * 8192 + t : nullptr;
  const int n_filters = BANK ? a.n_filters : 1;
  for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const float2 *prev = blk == 0 ? (const float2 *)a.hist_in : x + (size_t)(blk - 1) * kHalf;
    const float2 *cur = x + (size_t)blk * kHalf;
    {   // load, radix-2 DIF step, stage 1 of both halves
      float2 e[16], o[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) { e[q] = prev[256 * q + t]; o[q] = cur[256 * q + t]; }
      if (blk + (int)gridDim.x < n_blocks) {   // this CTA's next block pair: 64 KB into L2 while this one is computed
        const char *nx = (const char *)(cur + ((size_t)gridDim.x - 1) * kHalf);
        prefetch_l2(nx + 128 * t); prefetch_l2(nx + 128 * (t + 256));
      }
      if (blk == n_blocks - 1) {
        float2 *ho = (float2 *)a.hist_out;
#pragma unroll
        for (int q = 0; q < 16; ++q) ho[256 * q + t] = o[q];
      }
      const float2 wt = tab[kT8 + t];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float2 p = e[q], c = o[q];
        e[q] = caddf(p, c);
        o[q] = cmulf(csubf(p, c), root8k(wt, q));
      }
      dif_stage1<false>(e, o, H0, tab, t);
    }
    __syncthreads();
    dif_stage2<false>(H0, tab, t);
    __syncwarp();                // stage 3 reads what lanes of this warp wrote (see stage3_pos_conv)
    const int qb = t & 15, p0 = stage3_pos_conv(t);
    if (BANK) {      // forward stage 3 -> the CTA's scratch line
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float2 v[16];
        load16(v, h ? H1 : H0, p0);
        dft16<false>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) scratch[(h * 16 + q) * 256] = v[q];
      }
    }
    for (int f = 0; f < n_filters; ++f) {
      const float2 *kp = (const float2 *)a.kperm + (size_t)f * 8192 + t;
      // (stage 3,) spectrum multiply, first inverse stage (DFT16 over qc, twiddle conj w_256^(c qb)); the filter
      // spectrum is requested first so that its latency hides behind the shared-memory reads and the DFT
      {
        float2 v0[16], v1[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float2 *v = h ? v1 : v0;
          float2 k[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) k[q] = __ldg(kp + (h * 16 + q) * 256);
          if (BANK) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = scratch[(h * 16 + q) * 256];
          } else {
            load16(v, h ? H1 : H0, p0);
            dft16<false>(v);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = cmulf(v[q], k[q]);
#pragma unroll
          for (int q = 8; q < 16; ++q) v[q] = cmulf(v[q], __ldg(kp + (h * 16 + q) * 256));
          dft16<true>(v);
        }
#pragma unroll
        for (int c = 1; c < 16; ++c) {
          const float2 w = tab[kT2 + 16 * c + qb];
          v0[c] = cmulw<true>(v0[c], w);
          v1[c] = cmulw<true>(v1[c], w);
        }
        store16(v0, H0, p0);
        store16(v1, H1, p0);
      }
      __syncwarp();
      {   // second inverse stage: thread t = 16 qa + c, DFT16 over qb, twiddle conj w_4096^(qa (16 b + c))
        const int qa = t >> 4, c = t & 15;
        float2 *B = H0 + pos3(qa, 0, c);
        const float2 *T = tab + kT1 + 256 * qa + c;
        float2 e[16], o[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) { e[q] = B[pos3(0, q, 0)]; o[q] = B[pos3(0, q, 0) + kHalfPad]; }
        dft16<true>(e);
        dft16<true>(o);
#pragma unroll
        for (int b = 0; b < 16; ++b) {         // (w_4096^(qa (16 b + c)) is 1 only for qa = 0: no row to skip here)
          const float2 w = T[16 * b];
          B[pos3(0, b, 0)] = cmulw<true>(e[b], w);
          B[pos3(0, b, 0) + kHalfPad] = cmulw<true>(o[b], w);
        }
      }
      __syncthreads();
      {   // third inverse stage of both halves (thread t = 16 b + c, DFT16 over qa) and the overlap-save combine
        const float2 *B = H0 + pos3(0, t >> 4, t & 15);
        float2 z[16], v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) z[q] = B[pos3(q, 0, 0) + kHalfPad];
        dft16<true>(z);
        const float2 wt = tab[kT8 + t];
#pragma unroll
        for (int q = 0; q < 16; ++q) { z[q] = cmulw<true>(z[q], root8k(wt, q)); v[q] = B[pos3(q, 0, 0)]; }
        dft16<true>(v);
        float2 *o = (float2 *)a.out + (size_t)f * a.out_stride + (size_t)blk * kHalf + t;
#pragma unroll
        for (int q = 0; q < 16; ++q) o[256 * q] = csubf(v[q], z[q]);
      }
      if (BANK && f + 1 < n_filters) __syncthreads();   // the next filter's first-stage stores; the next BLOCK waits inside dif_stage1
    }
  }
}

int resident_ctas(const void *fn, int dev, std::atomic<int> *cache) {
  if (!cache[dev]) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kT, kSmemBytes);
    cache[dev] = sms * (per_sm > 0 ? per_sm : 1);
  }
  return cache[dev];
}

}  // namespace

// host: the shared tables, double precision rounded once
void fft8k_tables(std::vector<float> &tab) {
  tab.assign(2 * (size_t)kTabLen, 0.f);
  auto put = [&](int idx, double num, double den) {
    const double ang = -2.0 * M_PI * std::fmod(num, den) / den;
    tab[2 * idx] = (float)std::cos(ang); tab[2 * idx + 1] = (float)std::sin(ang);
  };
  for (int q = 0; q < 16; ++q) for (int u = 0; u < 256; ++u) put(kT1 + 256 * q + u, (double)(q * u), 4096.0);
  for (int q = 0; q < 16; ++q) for (int c = 0; c < 16; ++c) put(kT2 + 16 * q + c, (double)(q * c), 256.0);
  for (int u = 0; u < 256; ++u) put(kT8 + u, (double)u, 8192.0);
}

// host: spectrum K (8192 complex, natural order) -> the kernel's order, scaled by 1/8192 (exact: a power of two)
void fft8k_permute_kernel(const float *kern, float *kperm) {
  for (int h = 0; h < 2; ++h)
    for (int qc = 0; qc < 16; ++qc)
      for (int t = 0; t < 256; ++t) {
        const int m = (t >> 4) + 16 * (t & 15) + 256 * qc;
        const int src = 2 * m + h, dst = (h * 16 + qc) * 256 + t;
        kperm[2 * dst] = kern[2 * src] * (1.0f / 8192.0f);
        kperm[2 * dst + 1] = kern[2 * src + 1] * (1.0f / 8192.0f);
      }
}

int launch_fft8k(const void *in, void *out, int n, int inverse, size_t batch, const void *tab, cudaStream_t st) {
  if (batch == 0) return SDRG_OK;
  if (batch > 0x7fffffffull) return set_error(SDRG_ERR_ARG, "FFT plan: batch too large");
  const int dev = current_device();
  static std::atomic<int> attr[kMaxDevices];
  if (!attr[dev]) {
    SDRG_CUDA(cudaFuncSetAttribute(fft8k_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    SDRG_CUDA(cudaFuncSetAttribute(fft8k_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    SDRG_CUDA(cudaFuncSetAttribute(fft8k_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    SDRG_CUDA(cudaFuncSetAttribute(fft8k_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr[dev] = 1;
  }
  static std::atomic<int> res[kMaxDevices];
  const int resident = resident_ctas((const void *)fft8k_kernel<false, true>, dev, res);
  const size_t items = n == 8192 ? batch : (batch + 1) / 2;
  const unsigned grid = (unsigned)(items < (size_t)resident ? items : (size_t)resident);
  const float2 *i2 = (const float2 *)in; float2 *o2 = (float2 *)out; const float2 *t2 = (const float2 *)tab;
  if (n == 8192) {
    if (inverse) fft8k_kernel<true, true><<<grid, kT, kSmemBytes, st>>>(i2, o2, (int)batch, t2);
    else fft8k_kernel<false, true><<<grid, kT, kSmemBytes, st>>>(i2, o2, (int)batch, t2);
  } else {
    if (inverse) fft8k_kernel<true, false><<<grid, kT, kSmemBytes, st>>>(i2, o2, (int)batch, t2);
    else fft8k_kernel<false, false><<<grid, kT, kSmemBytes, st>>>(i2, o2, (int)batch, t2);
  }
  SDRG_CHECK_LAUNCH("fft8k_kernel");
  return SDRG_OK;
}

int conv8k_grid(size_t n_blocks) {      // CTAs launch_conv8k will use (the bank's scratch is 64 KB per CTA)
  const int dev = current_device();
  static std::atomic<int> attr[kMaxDevices];
  if (!attr[dev]) {
    if (cudaFuncSetAttribute(conv8k_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(conv8k_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    attr[dev] = 1;
  }
  static std::atomic<int> res[kMaxDevices];
  const int resident = resident_ctas((const void *)conv8k_kernel<true>, dev, res);
  return (int)(n_blocks < (size_t)resident ? n_blocks : (size_t)resident);
}

int launch_conv8k(const FilterArgs &a, size_t n_blocks, cudaStream_t st) {
  if (n_blocks == 0) return SDRG_OK;
  if (n_blocks > 0x7fffffffull) return set_error(SDRG_ERR_ARG, "FilterNode: too many blocks in one call");
  const int grid = conv8k_grid(n_blocks);
  if (grid <= 0) return set_error(SDRG_ERR_CUDA, "FilterNode: cannot configure the block-4096 kernel");
  if (a.n_filters > 1) {
    if (!a.spec) return set_error(SDRG_ERR_RUNTIME, "FilterNode: the filter bank needs its spectrum scratch");
    conv8k_kernel<true><<<grid, kT, kSmemBytes, st>>>(a, (int)n_blocks);
  } else {
    conv8k_kernel<false><<<grid, kT, kSmemBytes, st>>>(a, (int)n_blocks);
  }
  SDRG_CHECK_LAUNCH("conv8k_kernel");
  return SDRG_OK;
}

}  // namespace sdrg
