// iqbb_fold_perwin.cu -- IQBaseBand<float>, short windows (ss <= 256): one whole window per thread group.
//
// The per-sample weights of iqbb_fold_kernels.cu cost two table look-ups and two complex products per sample;
// for a window of a few dozen samples that set-up, not HBM, is the bound (0.20-0.37 of the roofline).  Regrouped
// by window, the same folded sum reads
//     out[w] = A(a_b) * sum_{j=0}^{ss+L-2} V(r_b, j) * x[n_b - (L-1) + j],     256 a_b + r_b = phase at n_b,
//     V(r, j) = sum_{d = max(0, j-L+1)}^{min(j, ss-1)} B(r, d) k[j-d],         B(r, d) = lut[((r + d inc) >> 8) & 127]
// (n_b = first sample of the window; this is SURVEY.md appendix C.5's g_m with the phase factored out exactly).
// B(r, d) = B0(d) * rho^carry(r, d) with carry(r, d) = [r + (d inc & 255) >= 256], so V(r, .) changes only where r
// crosses one of at most ss thresholds: there are <= ss + 1 DISTINCT rows, not 256, and the whole table
// ((ss+1)(ss+L-1) float2) fits into shared memory for the window lengths this kernel is for.  Per multiply-add:
// one LDS of x, one LDS of V, four FFMA -- (ss+L-1)/ss of them per input sample, no per-sample index arithmetic.
//
// Mapping.  A CTA stages a tile of TW = 256/G consecutive windows (<= 4096 samples + L-1 halo, read once from HBM,
// coalesced, prefetched into registers while the previous tile is summed); G = 1..16 threads share a window
// (thread g takes j = g, g+G, ...) and combine with shuffles.  Windows are laid out with a pitch of ss + delta
// samples, (ss + delta)/G odd, so that the G x 16/G lanes of a half warp hit 16 different banks.  A window whose
// halo is inside the call is summed and STORED by its thread group alone (nothing else contributes to its slot);
// the first and last windows of a call -- cut by the call boundary -- take the per-sample path of
// fold_chunk_general(), restricted to the (sample, window) pairs the whole windows do not cover.
#include "iqbb_fold_common.cuh"

namespace sdrg {
using namespace foldk;
namespace {

constexpr int kPwThreads = 256;
constexpr int kPwSpt = 16;                 // staged samples per thread and tile (tile <= 4096 samples + halo)
constexpr int kPwLoads = kPwSpt + 1;

struct PerwinGeom {
  uint32_t n_tiles, pitch, delta, magic;    // magic = ceil(2^32 / ss)
  uint32_t n_edge;                          // edge chunks: ids 0..d_lo-1 and d_hi..n_chunks-1
  uint32_t x_tile;                          // float2 elements of the staged tile
};

template <int G>
__device__ __forceinline__ void mac_run(float2 &a0, float2 &a1, const float2 *__restrict__ pv, const float2 *__restrict__ px, int cnt) {
  int t = 0;
  for (; t + 4 <= cnt; t += 4, pv += 4 * G, px += 4 * G) {
    const float2 v0 = pv[0], v1 = pv[G], v2 = pv[2 * G], v3 = pv[3 * G];
    const float2 x0 = px[0], x1 = px[G], x2 = px[2 * G], x3 = px[3 * G];
    cfma(a0, v0, x0); cfma(a1, v1, x1); cfma(a0, v2, x2); cfma(a1, v3, x3);
  }
  for (; t < cnt; ++t, pv += G, px += G) cfma(a0, pv[0], px[0]);
}

template <int LG>
__global__ void __launch_bounds__(kPwThreads, 3) iqbb_fold_f32_perwin_kernel(const IqbbFoldArgs a, const PerwinGeom geo) {
  constexpr int G = 1 << LG, TW = kPwThreads / G;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ float2 sA[128];
  __shared__ float2 sH[256];
  __shared__ uint16_t sC[256];
  float2 *sV = (float2 *)dyn_smem;
  float2 *sX = sV + (size_t)a.v_rows * a.v_pitch;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (uint32_t k = blockIdx.x * blockDim.x + tid; k < a.zero_next; k += gridDim.x * blockDim.x)
    ((float2 *)a.acc_next)[k] = make_float2(0.f, 0.f);
  if (tid < 128) sA[tid] = a.tab_a[tid];
  sH[tid] = a.tab_u[(size_t)tid * a.taps_len];
  sC[tid] = a.tab_cls[tid];
  for (uint32_t k = tid; k < a.v_rows * a.v_pitch; k += kPwThreads) sV[k] = a.tab_v[k];

  const float2 *__restrict__ x = (const float2 *)a.x;
  float *acc_out = (float *)a.acc_cur;
  const int ss = (int)a.ss, L1 = (int)a.taps_len - 1, len = ss + L1;
  const int win_off = (int)a.first - (int)a.r0;
  const int pitch = (int)geo.pitch, delta = (int)geo.delta;

  // tile t: windows s0 .. s0+nw-1, samples [s0*ss + win_off - L1, (s0+nw)*ss + win_off)
  float2 xv[kPwLoads];
  auto load_tile = [&](uint32_t t) {
    const uint32_t s0 = a.d_lo + t * TW;
    const int nw = (int)min((uint32_t)TW, a.d_hi + 1 - s0);
    const int count = nw * ss + L1;
    const float2 *__restrict__ xt = x + ((int64_t)s0 * ss + win_off - L1) + tid;
#pragma unroll
    for (int r = 0; r < kPwLoads; ++r) {
      xv[r] = make_float2(0.f, 0.f);
      if (tid + kPwThreads * r < count) xv[r] = ld_stream(xt + kPwThreads * r);
    }
  };
  __syncthreads();

  // Edge chunks, one per warp of the last CTA: windows before d_lo keep their own sums (their tails into d_lo are
  // part of that whole window), window d_hi only sends its tails ahead, everything behind it takes both.
  if (blockIdx.x == gridDim.x - 1) {
    const uint32_t inc32 = (32u * a.inc) & 0x7fffu, inc256 = (256u * a.inc) & 0x7fffu;
    WarpStage none{nullptr, 0u, 0u};
    for (uint32_t e = warp; e < geo.n_edge; e += kPwThreads / 32) {
      const uint32_t id = e < a.d_lo ? e : a.d_hi + (e - a.d_lo);
      const uint32_t what = e < a.d_lo ? (e + 1 < a.d_lo ? 3u : 1u) : (id == a.d_hi ? 2u : 3u);
      fold_chunk_general<false>(a, id, 1u, x, sA, sH, lane, L1, win_off, inc32, inc256, none, acc_out, what);
    }
  }

  if (blockIdx.x < geo.n_tiles) load_tile(blockIdx.x);

  // per-thread constants: its share of a window's ss+L-1 products, j = g, g+G, ... split at the halo / window border
  const int g = tid & (G - 1), wl = tid >> LG;
  const int nA = g < L1 ? (L1 - g + G - 1) / G : 0;
  const int jB0 = L1 + (((g - L1) % G) + G) % G;
  const int nB = jB0 < len ? (len - jB0 + G - 1) / G : 0;
  const float2 *px_h = sX + wl * pitch + g;                             // halo sample j at (B_w - delta - L1) + j
  const float2 *px_m = sX + (L1 + delta + wl * pitch) + (jB0 - L1);     // window sample j at B_w + (j - L1)

  for (uint32_t t = blockIdx.x; t < geo.n_tiles; t += gridDim.x) {
    const uint32_t s0 = a.d_lo + t * TW;
    const int nw = (int)min((uint32_t)TW, a.d_hi + 1 - s0);
    const int count = nw * ss + L1;
#pragma unroll
    for (int r = 0; r < kPwLoads; ++r) {
      const int m = tid + kPwThreads * r;
      if (m < count) sX[m + delta * (int)__umulhi((uint32_t)(m + ss - L1), geo.magic)] = xv[r];
    }
    __syncthreads();
    if (t + gridDim.x < geo.n_tiles) load_tile(t + gridDim.x);          // in flight while this tile is summed
    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
    uint32_t phb = 0;
    const bool active = wl < nw;
    if (active) {
      phb = a.phase0 + (uint32_t)((int)((s0 + wl) * a.ss) + win_off) * a.inc;        // bits 0..14: phase at the window's first sample
      const float2 *v = sV + (uint32_t)sC[phb & 255u] * a.v_pitch;
      mac_run<G>(a0, a1, v + g, px_h, nA);
      mac_run<G>(a0, a1, v + jB0, px_m, nB);
    }
    a0.x += a1.x; a0.y += a1.y;
#pragma unroll
    for (int d = G >> 1; d > 0; d >>= 1) { a0.x += __shfl_xor_sync(kFull, a0.x, d); a0.y += __shfl_xor_sync(kFull, a0.y, d); }
    if (active && g == 0)
      ((float2 *)acc_out)[s0 + wl] = cmul(sA[(phb & 0x7fffu) >> 8], a0);
    __syncthreads();
  }
}

template <int LG>
int launch_perwin_lg(const IqbbFoldArgs &a, const PerwinGeom &geo, size_t smem, cudaStream_t st) {
  // occupancy depends on the dynamic shared memory size, which depends on the configuration: cached per (device, size)
  static std::atomic<int> resident_dev[kMaxDevices], smem_for[kMaxDevices], smem_max[kMaxDevices];
  const int dev = current_device();
  if (smem_max[dev] < (int)smem) {
    SDRG_CUDA(cudaFuncSetAttribute(iqbb_fold_f32_perwin_kernel<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_max[dev] = (int)smem;
  }
  if (smem_for[dev] != (int)smem || !resident_dev[dev]) {
    int sms = 0, per_sm = 0;
    SDRG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SDRG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, iqbb_fold_f32_perwin_kernel<LG>, kPwThreads, smem));
    resident_dev[dev] = sms * (per_sm > 0 ? per_sm : 1);
    smem_for[dev] = (int)smem;
  }
  const uint64_t resident = (uint64_t)resident_dev[dev];
  const uint64_t want = geo.n_tiles > 0 ? geo.n_tiles : 1;
  iqbb_fold_f32_perwin_kernel<LG><<<(unsigned)(want < resident ? want : resident), kPwThreads, smem, st>>>(a, geo);
  SDRG_CHECK_LAUNCH("iqbb_fold_f32_perwin_kernel");
  return SDRG_OK;
}

int lg_of(uint32_t ss) {          // threads per window: the tile (256/G windows) stays within 4096 samples
  int lg = 0;
  while ((256u >> lg) * ss > (uint32_t)(kPwThreads * kPwSpt)) ++lg;
  return lg;
}

constexpr size_t kPwMaxSmem = 200 * 1024;

}  // namespace

// Fills in d_lo / d_hi; true when the call has whole windows and the table + tile fit into shared memory.
bool fold_perwin_eligible(IqbbFoldArgs &a) {
  if (!a.tab_v || !a.tab_cls || a.v_rows == 0 || a.ss < 2 || a.ss > 256 || a.taps_len > a.ss + 1) return false;
  const int64_t ss = a.ss, L1 = (int64_t)a.taps_len - 1, win_off = (int64_t)a.first - (int64_t)a.r0;
  int64_t lo = 1;
  while (lo * ss + win_off - L1 < 0) ++lo;
  const int64_t hi = ((int64_t)a.n - win_off) / ss - 1;        // (hi + 1) ss + win_off <= n
  if (hi < lo) return false;
  a.d_lo = (uint32_t)lo; a.d_hi = (uint32_t)hi;
  const int lg = lg_of(a.ss);
  if (lg > 4) return false;
  const size_t tile = (size_t)(L1 + 2 * 16 + (256 >> lg) * (ss + 2 * 16));
  return ((size_t)a.v_rows * a.v_pitch + tile) * sizeof(float2) <= kPwMaxSmem;
}

int launch_fold_perwin(const IqbbFoldArgs &a_in, cudaStream_t st) {
  IqbbFoldArgs a = a_in;
  const int lg = lg_of(a.ss), G = 1 << lg;
  const uint32_t L1 = a.taps_len - 1;
  const uint64_t q_last = (uint64_t)a.r0 + (a.n - 1) - ((a.first && a.n > 1) ? 1 : 0);
  a.part = 8192; a.cpw = 1; a.pf_dist = 0; a.fast = 0;
  a.n_chunks = (uint32_t)(q_last / a.ss + 1);
  PerwinGeom geo{};
  uint32_t delta = 0;
  while ((a.ss + delta) % G != 0 || (((a.ss + delta) / G) & 1u) == 0) ++delta;     // (ss + delta) / G odd
  geo.delta = delta; geo.pitch = a.ss + delta;
  geo.magic = (uint32_t)(((1ull << 32) + a.ss - 1) / a.ss);
  const uint32_t tw = 256u >> lg;
  geo.n_tiles = (a.d_hi - a.d_lo + 1 + tw - 1) / tw;
  geo.n_edge = a.d_lo + (a.n_chunks > a.d_hi ? a.n_chunks - a.d_hi : 0u);
  geo.x_tile = L1 + delta + tw * geo.pitch;
  const size_t smem = ((size_t)a.v_rows * a.v_pitch + geo.x_tile) * sizeof(float2);
  switch (lg) {
    case 0: return launch_perwin_lg<0>(a, geo, smem, st);
    case 1: return launch_perwin_lg<1>(a, geo, smem, st);
    case 2: return launch_perwin_lg<2>(a, geo, smem, st);
    case 3: return launch_perwin_lg<3>(a, geo, smem, st);
    case 4: return launch_perwin_lg<4>(a, geo, smem, st);
    default: return set_error(SDRG_ERR_RUNTIME, "IQBaseBand<float>: per-window kernel needs sub_sample <= 256");
  }
}

}  // namespace sdrg
