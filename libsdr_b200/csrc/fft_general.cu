// fft_general.cu -- FFTPlan<float> for ANY size and the FilterNode path for ANY block size.
//
// The reference hands every size to FFTW (src/fftplan_fftw3.hh:83-106: fftwf_plan_dft_1d(N, ...) for whatever N the
// buffer has; FilterNode(block_size) accepts any block, src/filternode.hh:236).  The shared-memory kernels
// (fft_kernels.cu, fft8k_kernels.cu) cover the powers of two up to 8192 -- the hot sizes.  Everything else is built
// here out of them, trading extra passes over HBM for generality (cold-path sizes):
//   * powers of two above 8192: the four-step decomposition n = n1 n2 (both <= 8192): transpose, n2 row FFTs of n1
//     points, twiddle w_n^(j2 k1) + transpose, n1 row FFTs of n2 points, transpose.  The twiddle is a product of two
//     table entries (w_n^(8192 hi) w_n^lo), built in double on the host.
//   * every other size: Bluestein's chirp-z identity j k = (j^2 + k^2 - (k - j)^2) / 2 turns the n-point DFT into a
//     circular convolution of length M = 2^ceil(log2(2n - 1)), run on the power-of-two engine; chirps in double with
//     the phase index j^2 reduced mod 2n in integers.
//   * FilterNode blocks that are not a power of two or exceed 4096: overlap-save with FFT size M = 2^ceil(log2 2B)
//     (segments [last M - B samples | B new samples], the last B outputs of each are the block's result) -- the same
//     causal linear convolution with the same taps and the same normalisation as the reference's 2B-point transform
//     pair: gather segments, batched FFT_M, per filter multiply + inverse FFT_M + crop.
#include "fft_kernels.cuh"

#include <cmath>
#include <complex>
#include <vector>

namespace sdrg {
namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// out[b][c][r] = in[b][r][c] * (tw ? w_n^(r c) : 1), matrices rows x cols per batch entry, 32 x 32 tiles
__global__ void transpose_tw_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int rows, int cols,
                                    const float2 *__restrict__ tw_hi, const float2 *__restrict__ tw_lo, int inverse) {
  __shared__ float2 tile[32][33];
  const size_t mat = (size_t)rows * cols;
  const float2 *src = in + blockIdx.z * mat;
  float2 *dst = out + blockIdx.z * mat;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const int r = r0 + k, c = c0 + threadIdx.x;
    if (r < rows && c < cols) {
      float2 v = src[(size_t)r * cols + c];
      if (tw_hi) {
        const uint64_t m = (uint64_t)r * (uint64_t)c;                     // < n
        float2 w = cmul(tw_hi[m >> 13], tw_lo[m & 8191]);
        if (inverse) w.y = -w.y;
        v = cmul(v, w);
      }
      tile[k][threadIdx.x] = v;
    }
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const int c = c0 + k, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][k];
  }
}

// Bluestein, before the convolution: a[b][j] = (conj?) x[b][j] * c[j] for j < n, 0 up to M
__global__ void bluestein_pre_kernel(const float2 *__restrict__ x, float2 *__restrict__ a, const float2 *__restrict__ chirp,
                                     int n, int M, int inverse) {
  const size_t b = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (j < n) { v = x[b * n + j]; if (inverse) v.y = -v.y; v = cmul(v, chirp[j]); }
    a[b * M + j] = v;
  }
}
// a[b][k] *= g[k]  (spectrum of the chirp kernel, or a filter spectrum)
__global__ void pointwise_mul_kernel(float2 *__restrict__ a, const float2 *__restrict__ g, int M) {
  const size_t b = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x) a[b * M + k] = cmul(a[b * M + k], g[k]);
}
__global__ void pointwise_mul_to_kernel(const float2 *__restrict__ a, const float2 *__restrict__ g, float2 *__restrict__ o, int M) {
  const size_t b = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x) o[b * M + k] = cmul(a[b * M + k], g[k]);
}
// after: X[b][k] = (conj?)(c[k] * conv[b][k] / M)
__global__ void bluestein_post_kernel(const float2 *__restrict__ conv, float2 *__restrict__ X, const float2 *__restrict__ chirp,
                                      int n, int M, int inverse) {
  const size_t b = blockIdx.y;
  const float sc = 1.0f / (float)M;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    float2 v = cmul(conv[b * M + k], chirp[k]);
    v.x *= sc; v.y *= sc;
    if (inverse) v.y = -v.y;
    X[b * n + k] = v;
  }
}

// FilterNode, general path: segment s = [the M - B samples before block s | block s]; samples before the call come
// from hist (the last M - B samples of the stream so far, oldest first)
__global__ void ola_gather_kernel(const float2 *__restrict__ x, const float2 *__restrict__ hist, float2 *__restrict__ seg,
                                  int B, int M, long long s0) {
  const long long s = blockIdx.y;
  const int Hh = M - B;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x) {
    const long long i = (s0 + s) * B - Hh + k;                           // call-relative sample index
    seg[s * M + k] = i >= 0 ? x[i] : hist[Hh + i];
  }
}
__global__ void ola_crop_kernel(const float2 *__restrict__ y, float2 *__restrict__ out, int B, int M, float scale) {
  const long long s = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < B; k += gridDim.x * blockDim.x) {
    const float2 v = y[s * M + (M - B) + k];
    out[s * B + k] = make_float2(v.x * scale, v.y * scale);
  }
}
// new history = last Hh samples of [hist | x(n)]
__global__ void ola_roll_kernel(const float2 *__restrict__ x, long long n, const float2 *__restrict__ hist, float2 *__restrict__ hist_out, int Hh) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Hh; k += gridDim.x * blockDim.x) {
    const long long i = n - Hh + k;
    hist_out[k] = i >= 0 ? x[i] : hist[Hh + i];
  }
}

int ilog2c(size_t n) { int l = 0; while (((size_t)1 << l) < n) ++l; return l; }
bool is_pow2(size_t n) { return n && !(n & (n - 1)); }

int upload(const std::vector<float> &v, void **d) {
  SDRG_CUDA(cudaMalloc(d, v.size() * sizeof(float)));
  SDRG_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return SDRG_OK;
}
void roots(std::vector<float> &v, size_t count, double num_step, double den) {   // exp(-2 pi i k num_step / den), k < count
  v.resize(2 * count);
  for (size_t k = 0; k < count; ++k) {
    const double a = -2.0 * M_PI * std::fmod((double)k * num_step, den) / den;
    v[2 * k] = (float)std::cos(a); v[2 * k + 1] = (float)std::sin(a);
  }
}

int grow(void **p, size_t *cap, size_t need) {
  if (*cap >= need && *p) return SDRG_OK;
  if (*p) { SDRG_CUDA(cudaDeviceSynchronize()); SDRG_CUDA(cudaFree(*p)); }
  *p = nullptr; *cap = 0;
  SDRG_CUDA(cudaMalloc(p, need ? need : 16));
  *cap = need;
  return SDRG_OK;
}

void host_fft_pow2(std::vector<std::complex<double> > &a) {
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1)
    for (size_t k = 0; k < len / 2; ++k) {
      const std::complex<double> w = std::polar(1.0, -2.0 * M_PI * (double)k / (double)len);
      for (size_t s = 0; s < n; s += len) {
        const std::complex<double> u = a[s + k], t = a[s + k + len / 2] * w;
        a[s + k] = u + t; a[s + k + len / 2] = u - t;
      }
    }
}

}  // namespace

// ---- power-of-two engine: any n = 2^L >= 2 ------------------------------------------------------------------------
Pow2Fft::~Pow2Fft() {
  cudaFree(d_tw); cudaFree(d_tab8k); cudaFree(d_tw_hi); cudaFree(d_tw_lo); cudaFree(d_s0); cudaFree(d_s1);
  delete sub1; delete sub2;
}
int Pow2Fft::init(size_t n_) {
  n = n_;
  if (!is_pow2(n) || n < 2) return set_error(SDRG_ERR_ARG, "Pow2Fft: %zu is not a power of two", n);
  log2n = ilog2c(n);
  std::vector<float> v;
  if (n <= 8192) {
    roots(v, n, 1.0, (double)n);
    int rc = upload(v, &d_tw);
    if (rc) return rc;
    if (n == 4096 || n == 8192) { fft8k_tables(v); rc = upload(v, &d_tab8k); if (rc) return rc; }
    return SDRG_OK;
  }
  const int l1 = log2n / 2, l2 = log2n - l1;
  n1 = (size_t)1 << l1; n2 = (size_t)1 << l2;
  if (n2 > 8192) return set_error(SDRG_ERR_CONFIG, "FFT plan: size %zu exceeds the supported maximum 2^26", n);
  roots(v, (n + 8191) / 8192, 8192.0, (double)n);
  int rc = upload(v, &d_tw_hi);
  if (rc) return rc;
  roots(v, 8192, 1.0, (double)n);
  if ((rc = upload(v, &d_tw_lo))) return rc;
  sub1 = new Pow2Fft(); sub2 = new Pow2Fft();
  if ((rc = sub1->init(n1))) return rc;
  return sub2->init(n2);
}
int Pow2Fft::exec(const void *in, void *out, size_t batch, int inverse, cudaStream_t st) {
  if (!batch) return SDRG_OK;
  if (n <= 8192) {
    if (d_tab8k) return launch_fft8k(in, out, (int)n, inverse, batch, d_tab8k, st);
    return launch_fft_batch(in, out, (int)n, log2n, inverse, batch, d_tw, st);
  }
  if (batch > 65535) return set_error(SDRG_ERR_ARG, "FFT plan: at most 65535 transforms of %zu points per call", n);
  const size_t bytes = batch * n * sizeof(float2);
  int rc = grow(&d_s0, &cap0, bytes);
  if (rc) return rc;
  if ((rc = grow(&d_s1, &cap1, bytes))) return rc;
  const dim3 thr(32, 8);
  // x as [n1][n2] -> T[n2][n1]
  transpose_tw_kernel<<<dim3((unsigned)((n2 + 31) / 32), (unsigned)((n1 + 31) / 32), (unsigned)batch), thr, 0, st>>>(
      (const float2 *)in, (float2 *)d_s0, (int)n1, (int)n2, nullptr, nullptr, 0);
  SDRG_CHECK_LAUNCH("transpose_tw_kernel");
  if ((rc = sub1->exec(d_s0, d_s1, batch * n2, inverse, st))) return rc;          // rows of n1 points -> [j2][k1]
  // twiddle w_n^(j2 k1), -> U[k1][j2]
  transpose_tw_kernel<<<dim3((unsigned)((n1 + 31) / 32), (unsigned)((n2 + 31) / 32), (unsigned)batch), thr, 0, st>>>(
      (const float2 *)d_s1, (float2 *)d_s0, (int)n2, (int)n1, (const float2 *)d_tw_hi, (const float2 *)d_tw_lo, inverse);
  SDRG_CHECK_LAUNCH("transpose_tw_kernel");
  if ((rc = sub2->exec(d_s0, d_s1, batch * n1, inverse, st))) return rc;          // rows of n2 points -> [k1][k2]
  // X[k1 + n1 k2]: -> [k2][k1]
  transpose_tw_kernel<<<dim3((unsigned)((n2 + 31) / 32), (unsigned)((n1 + 31) / 32), (unsigned)batch), thr, 0, st>>>(
      (const float2 *)d_s1, (float2 *)out, (int)n1, (int)n2, nullptr, nullptr, 0);
  SDRG_CHECK_LAUNCH("transpose_tw_kernel");
  return SDRG_OK;
}

// ---- any size ---------------------------------------------------------------------------------------------------------
AnyFft::~AnyFft() { cudaFree(d_chirp); cudaFree(d_ghat); cudaFree(d_a); delete p2; }
int AnyFft::init(size_t n_) {
  n = n_;
  if (n < 1) return set_error(SDRG_ERR_CONFIG, "Can not construct FFT plan: Buffer is empty!");
  if (n > ((size_t)1 << 24)) return set_error(SDRG_ERR_CONFIG, "FFT plan: size %zu exceeds the supported maximum 2^24", n);
  p2 = new Pow2Fft();
  if (is_pow2(n) && n >= 2) { M = n; return p2->init(n); }
  if (n == 1) { M = 1; return SDRG_OK; }
  M = (size_t)1 << ilog2c(2 * n - 1);
  int rc = p2->init(M);
  if (rc) return rc;
  // chirp c[j] = exp(-i pi j^2 / n), phase index j^2 mod 2n kept exact in integers
  std::vector<float> c(2 * n);
  std::vector<std::complex<double> > g(M, std::complex<double>(0, 0));
  for (size_t j = 0; j < n; ++j) {
    const uint64_t q = ((uint64_t)j * (uint64_t)j) % (2 * (uint64_t)n);
    const double ph = M_PI * (double)q / (double)n;
    c[2 * j] = (float)std::cos(ph); c[2 * j + 1] = (float)(-std::sin(ph));
    const std::complex<double> gj(std::cos(ph), std::sin(ph));          // conj(c[j])
    g[j] = gj;
    if (j) g[M - j] = gj;
  }
  host_fft_pow2(g);
  std::vector<float> gh(2 * M);
  for (size_t k = 0; k < M; ++k) { gh[2 * k] = (float)g[k].real(); gh[2 * k + 1] = (float)g[k].imag(); }
  if ((rc = upload(c, &d_chirp))) return rc;
  return upload(gh, &d_ghat);
}
int AnyFft::exec(const void *in, void *out, size_t batch, int inverse, cudaStream_t st) {
  if (!batch) return SDRG_OK;
  if (n == 1) { if (in != out) SDRG_CUDA(cudaMemcpyAsync(out, in, batch * sizeof(float2), cudaMemcpyDeviceToDevice, st)); return SDRG_OK; }
  if (!d_chirp) return p2->exec(in, out, batch, inverse, st);
  // Bluestein; the batch is cut so that the scratch stays below 1 GiB
  const size_t per = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(batch, 65535), ((size_t)1 << 30) / (M * sizeof(float2))));
  int rc = grow(&d_a, &cap_a, per * M * sizeof(float2));
  if (rc) return rc;
  for (size_t b0 = 0; b0 < batch; b0 += per) {
    const size_t nb = std::min(per, batch - b0);
    const float2 *x = (const float2 *)in + b0 * n;
    float2 *X = (float2 *)out + b0 * n, *a = (float2 *)d_a;
    const dim3 g((unsigned)std::min<size_t>((M + 255) / 256, 1024), (unsigned)nb);
    bluestein_pre_kernel<<<g, 256, 0, st>>>(x, a, (const float2 *)d_chirp, (int)n, (int)M, inverse);
    SDRG_CHECK_LAUNCH("bluestein_pre_kernel");
    if ((rc = p2->exec(a, a, nb, 0, st))) return rc;
    pointwise_mul_kernel<<<g, 256, 0, st>>>(a, (const float2 *)d_ghat, (int)M);
    SDRG_CHECK_LAUNCH("pointwise_mul_kernel");
    if ((rc = p2->exec(a, a, nb, 1, st))) return rc;
    bluestein_post_kernel<<<g, 256, 0, st>>>(a, X, (const float2 *)d_chirp, (int)n, (int)M, inverse);
    SDRG_CHECK_LAUNCH("bluestein_post_kernel");
  }
  return SDRG_OK;
}

// ---- FilterNode, general block sizes ----------------------------------------------------------------------------------
GeneralOla::~GeneralOla() { cudaFree(d_seg); cudaFree(d_work); delete p2; }
int GeneralOla::init(size_t block) {
  B = block;
  M = (size_t)1 << ilog2c(2 * B);
  p2 = new Pow2Fft();
  return p2->init(M);
}
// kern: n_filters spectra of M points (already divided by the reference's norm); hist: M - B samples
int GeneralOla::run(const void *x, size_t n_blocks, const void *hist_in, void *hist_out, const void *kern, int n_filters,
                    void *out, size_t out_stride, cudaStream_t st) {
  if (!n_blocks) return SDRG_OK;
  const size_t per = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(n_blocks, 65535), ((size_t)1 << 29) / (M * sizeof(float2))));
  int rc = grow(&d_seg, &cap_seg, per * M * sizeof(float2));
  if (rc) return rc;
  if ((rc = grow(&d_work, &cap_work, per * M * sizeof(float2)))) return rc;
  const int Hh = (int)(M - B);
  for (size_t s0 = 0; s0 < n_blocks; s0 += per) {
    const size_t ns = std::min(per, n_blocks - s0);
    const dim3 g((unsigned)std::min<size_t>((M + 255) / 256, 1024), (unsigned)ns);
    ola_gather_kernel<<<g, 256, 0, st>>>((const float2 *)x, (const float2 *)hist_in, (float2 *)d_seg, (int)B, (int)M, (long long)s0);
    SDRG_CHECK_LAUNCH("ola_gather_kernel");
    if ((rc = p2->exec(d_seg, d_seg, ns, 0, st))) return rc;
    for (int f = 0; f < n_filters; ++f) {
      pointwise_mul_to_kernel<<<g, 256, 0, st>>>((const float2 *)d_seg, (const float2 *)kern + (size_t)f * M, (float2 *)d_work, (int)M);
      SDRG_CHECK_LAUNCH("pointwise_mul_to_kernel");
      if ((rc = p2->exec(d_work, d_work, ns, 1, st))) return rc;
      ola_crop_kernel<<<g, 256, 0, st>>>((const float2 *)d_work, (float2 *)out + (size_t)f * out_stride + s0 * B, (int)B, (int)M, 1.0f / (float)M);
      SDRG_CHECK_LAUNCH("ola_crop_kernel");
    }
  }
  ola_roll_kernel<<<(unsigned)std::min<size_t>((Hh + 255) / 256 + 1, 1024), 256, 0, st>>>((const float2 *)x, (long long)(n_blocks * B),
                                                                                         (const float2 *)hist_in, (float2 *)hist_out, Hh);
  SDRG_CHECK_LAUNCH("ola_roll_kernel");
  return SDRG_OK;
}


// ---- FFTPlan<double> (src/fftplan_fftw3.hh:12-75) -----------------------------------------------------------------------
// Not on the hot path (nothing in the receive chain transforms doubles): radix-2 Stockham passes through global memory
// for powers of two, Bluestein on top of them for every other size, all arithmetic in double on the device.
namespace {
__device__ __forceinline__ double2 cmuld(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void radix2_stage_f64(const double2 *__restrict__ in, double2 *__restrict__ out, unsigned n, unsigned Ns,
                                 const double2 *__restrict__ tw, int inverse) {
  const size_t b = blockIdx.y;
  const unsigned half = n >> 1, tstep = half / Ns;
  for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < half; j += gridDim.x * blockDim.x) {
    const unsigned k = j & (Ns - 1);
    double2 w = tw[k * tstep];                       // exp(-2 pi i k / (2 Ns))
    if (inverse) w.y = -w.y;
    const double2 a = in[b * n + j], c = cmuld(in[b * n + j + half], w);
    const size_t o = b * n + (size_t)(j - k) * 2 + k;
    out[o] = make_double2(a.x + c.x, a.y + c.y);
    out[o + Ns] = make_double2(a.x - c.x, a.y - c.y);
  }
}
__global__ void bluestein_pre_f64(const double2 *__restrict__ x, double2 *__restrict__ a, const double2 *__restrict__ chirp, int n, int M, int inverse) {
  const size_t b = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
    double2 v = make_double2(0.0, 0.0);
    if (j < n) { v = x[b * n + j]; if (inverse) v.y = -v.y; v = cmuld(v, chirp[j]); }
    a[b * M + j] = v;
  }
}
__global__ void pointwise_mul_f64(double2 *__restrict__ a, const double2 *__restrict__ g, int M) {
  const size_t b = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x) a[b * M + k] = cmuld(a[b * M + k], g[k]);
}
__global__ void bluestein_post_f64(const double2 *__restrict__ conv, double2 *__restrict__ X, const double2 *__restrict__ chirp, int n, int M, int inverse) {
  const size_t b = blockIdx.y;
  const double sc = 1.0 / (double)M;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double2 v = cmuld(conv[b * M + k], chirp[k]);
    v.x *= sc; v.y *= sc;
    if (inverse) v.y = -v.y;
    X[b * n + k] = v;
  }
}
}  // namespace

Fft64::~Fft64() { cudaFree(d_tw); cudaFree(d_chirp); cudaFree(d_ghat); cudaFree(d_a); cudaFree(d_b); }
int Fft64::init(size_t n_) {
  n = n_;
  if (n < 1) return set_error(SDRG_ERR_CONFIG, "Can not construct FFT plan: Buffer is empty!");
  if (n > ((size_t)1 << 22)) return set_error(SDRG_ERR_CONFIG, "FFT plan (double): size %zu exceeds the supported maximum 2^22", n);
  M = is_pow2(n) ? n : (size_t)1 << ilog2c(2 * n - 1);
  std::vector<double> tw(M >= 2 ? M : 2);
  for (size_t k = 0; k < M / 2; ++k) { const double a = -2.0 * M_PI * (double)k / (double)M; tw[2 * k] = std::cos(a); tw[2 * k + 1] = std::sin(a); }
  SDRG_CUDA(cudaMalloc(&d_tw, tw.size() * sizeof(double)));
  SDRG_CUDA(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (M == n) return SDRG_OK;
  std::vector<double> c(2 * n), gh(2 * M);
  std::vector<std::complex<double> > g(M, std::complex<double>(0, 0));
  for (size_t j = 0; j < n; ++j) {
    const uint64_t q = ((uint64_t)j * (uint64_t)j) % (2 * (uint64_t)n);
    const double ph = M_PI * (double)q / (double)n;
    c[2 * j] = std::cos(ph); c[2 * j + 1] = -std::sin(ph);
    g[j] = std::complex<double>(std::cos(ph), std::sin(ph));
    if (j) g[M - j] = g[j];
  }
  host_fft_pow2(g);
  for (size_t k = 0; k < M; ++k) { gh[2 * k] = g[k].real(); gh[2 * k + 1] = g[k].imag(); }
  SDRG_CUDA(cudaMalloc(&d_chirp, c.size() * sizeof(double)));
  SDRG_CUDA(cudaMemcpy(d_chirp, c.data(), c.size() * sizeof(double), cudaMemcpyHostToDevice));
  SDRG_CUDA(cudaMalloc(&d_ghat, gh.size() * sizeof(double)));
  SDRG_CUDA(cudaMemcpy(d_ghat, gh.data(), gh.size() * sizeof(double), cudaMemcpyHostToDevice));
  return SDRG_OK;
}
// M-point transform of `batch` rows held in d_a; the result ends up in d_a again
int Fft64::pow2_inplace(size_t batch, int inverse, cudaStream_t st) {
  double2 *src = (double2 *)d_a, *dst = (double2 *)d_b;
  const dim3 g((unsigned)std::min<size_t>((M / 2 + 255) / 256 + 1, 2048), (unsigned)batch);
  for (size_t Ns = 1; Ns < M; Ns <<= 1) {
    radix2_stage_f64<<<g, 256, 0, st>>>(src, dst, (unsigned)M, (unsigned)Ns, (const double2 *)d_tw, inverse);
    SDRG_CHECK_LAUNCH("radix2_stage_f64");
    std::swap(src, dst);
  }
  if (src != (double2 *)d_a) SDRG_CUDA(cudaMemcpyAsync(d_a, src, batch * M * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  return SDRG_OK;
}
int Fft64::exec(const void *in, void *out, size_t batch, int inverse, cudaStream_t st) {
  if (!batch) return SDRG_OK;
  const size_t per = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(batch, 65535), ((size_t)1 << 29) / (M * sizeof(double2))));
  int rc = grow(&d_a, &cap_a, per * M * sizeof(double2));
  if (rc) return rc;
  if ((rc = grow(&d_b, &cap_b, per * M * sizeof(double2)))) return rc;
  for (size_t b0 = 0; b0 < batch; b0 += per) {
    const size_t nb = std::min(per, batch - b0);
    const double2 *x = (const double2 *)in + b0 * n;
    double2 *X = (double2 *)out + b0 * n;
    const dim3 g((unsigned)std::min<size_t>((M + 255) / 256, 1024), (unsigned)nb);
    if (M == n) {
      SDRG_CUDA(cudaMemcpyAsync(d_a, x, nb * n * sizeof(double2), cudaMemcpyDeviceToDevice, st));
      if ((rc = pow2_inplace(nb, inverse, st))) return rc;
      SDRG_CUDA(cudaMemcpyAsync(X, d_a, nb * n * sizeof(double2), cudaMemcpyDeviceToDevice, st));
      continue;
    }
    bluestein_pre_f64<<<g, 256, 0, st>>>(x, (double2 *)d_a, (const double2 *)d_chirp, (int)n, (int)M, inverse);
    SDRG_CHECK_LAUNCH("bluestein_pre_f64");
    if ((rc = pow2_inplace(nb, 0, st))) return rc;
    pointwise_mul_f64<<<g, 256, 0, st>>>((double2 *)d_a, (const double2 *)d_ghat, (int)M);
    SDRG_CHECK_LAUNCH("pointwise_mul_f64");
    if ((rc = pow2_inplace(nb, 1, st))) return rc;
    bluestein_post_f64<<<g, 256, 0, st>>>((const double2 *)d_a, X, (const double2 *)d_chirp, (int)n, (int)M, inverse);
    SDRG_CHECK_LAUNCH("bluestein_post_f64");
  }
  return SDRG_OK;
}

}  // namespace sdrg
