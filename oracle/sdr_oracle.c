/* sdr_oracle.c -- CPU restatement of libsdr's receive-chain hot path (see sdr_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: checker / reported CPU baseline, never the product path.
 * Build with -fwrapv -fno-fast-math and WITHOUT -march=native (no FMA contraction), so that the
 * double-precision design formulas round exactly like the reference built with its release flags
 * (-O3, /root/reference/CMakeLists.txt:48-52).
 *
 * Every routine is sample-serial with an explicit state struct; signed overflow is two's
 * complement wrap (the reference's observable contract).
 */
#include "sdr_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- wrap helpers ---------------------------------------------------------------------------- */
static inline int32_t w32_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t w32_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t w32_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t asr32(int32_t a, int s) { return a >> s; }  /* gcc: arithmetic shift */

/* ---- IQBaseBand ------------------------------------------------------------------------------ */

/* FreqShiftBase ctor, src/freqshift.hh:26-36: lut[j] = 2^shift * exp(-2 pi i j/128), assigned to
 * complex<SScalar> (truncation toward zero).  shift: int16 -> 16, int8 -> 8, float -> 0
 * (src/traits.cc:11-29). */
static void build_lut(orc_iqbb *s) {
  int shift = (s->scalar == ORC_S16) ? 16 : (s->scalar == ORC_S8 ? 8 : 0);
  double scale = (double)(1 << shift);
  for (size_t i = 0; i < ORC_LUT_SIZE; i++) {
    double complex e = cexp(CMPLX(0.0, -(2 * M_PI * i) / ORC_LUT_SIZE));
    double re = scale * creal(e), im = scale * cimag(e);
    s->lutd_r[i] = re; s->lutd_i[i] = im;
    if (s->scalar == ORC_S8) {       /* complex<int16_t> LUT */
      s->lut_r[i] = (int16_t)re; s->lut_i[i] = (int16_t)im;
    } else {                         /* complex<int32_t> LUT */
      s->lut_r[i] = (int32_t)re; s->lut_i[i] = (int32_t)im;
    }
  }
}

void orc_iqbb_init(orc_iqbb *s, int scalar, double Fc, double Ff, double width,
                   size_t order, size_t sub_sample, double oFs) {
  memset(s, 0, sizeof(*s));
  s->scalar = scalar;
  s->freq_shift = Fc;                 /* FreqShiftBase<Scalar>(Fc, 0): double kept as is */
  s->Fc = (int32_t)Fc; s->Ff = (int32_t)Ff; s->Fs = 0; s->width = (int32_t)width;
  s->order = order < 1 ? 1 : order;   /* std::max(size_t(1), order) */
  if (s->order > ORC_MAX_ORDER) s->order = ORC_MAX_ORDER;
  s->sub_sample = sub_sample;
  s->oFs = oFs;
  build_lut(s);
}

/* src/freqshift.hh:78-87 */
static void update_lut_incr(orc_iqbb *s) {
  s->lut_inc = (size_t)((ORC_LUT_SIZE * (1 << 8) * fabs(s->freq_shift)) / s->nco_Fs);
  s->lut_count = 0;
}

/* src/baseband.hh:239-262 */
static void update_filter_kernel(orc_iqbb *s) {
  static double ar[ORC_MAX_ORDER], ai[ORC_MAX_ORDER];
  double w = (M_PI * s->width) / (s->Fs);
  double M = (double)(s->order) / 2.;
  double norm = 0;
  for (size_t i = 0; i < s->order; i++) {
    double a;
    if (s->order == 2 * i) { a = 4 * (w / M_PI); }
    else { a = sin(w * (i - M)) / (w * (i - M)); }
    double complex e = cexp(CMPLX(0.0, (-2 * M_PI * s->Ff * i) / s->Fs));
    /* complex<double>(a,0) *= e : (a*c - 0*d, a*d + 0*c) */
    double re = a * creal(e) - 0.0 * cimag(e);
    double im = a * cimag(e) + 0.0 * creal(e);
    double win = (0.42 - 0.5 * cos((2 * M_PI * i) / s->order) + 0.08 * cos((4 * M_PI * i) / s->order));
    re *= win; im *= win;
    ar[i] = re; ai[i] = im;
    norm += hypot(re, im);            /* std::abs(complex<double>) */
  }
  for (size_t i = 0; i < s->order; i++) {
    double re = ((double)(1 << 14) * ar[i]) / norm;
    double im = ((double)(1 << 14) * ai[i]) / norm;
    s->kr[i] = (int32_t)re; s->ki[i] = (int32_t)im;   /* complex<double> -> complex<int32_t> */
    s->kdr[i] = ar[i] / norm; s->kdi[i] = ai[i] / norm; /* float variant: no 2^14 */
  }
}

void orc_iqbb_set_center_frequency(orc_iqbb *s, double Fc) {
  s->Fc = (int32_t)Fc;                /* _Fc = Fc (int32 member) */
  s->freq_shift = (double)s->Fc;      /* setFrequencyShift(_Fc): the truncated value */
  if (s->nco_Fs != 0) update_lut_incr(s); else { s->lut_inc = 0; s->lut_count = 0; }
}

void orc_iqbb_set_filter_frequency(orc_iqbb *s, double Ff) {
  s->Ff = (int32_t)Ff;
  if (s->Fs != 0) update_filter_kernel(s);
}

int orc_iqbb_config(orc_iqbb *s, double sample_rate, size_t buffer_size) {
  if (sample_rate == 0 || buffer_size == 0) return 1;  /* incomplete config: silently ignored */
  s->Fs = (int32_t)sample_rate;
  s->source_bs = buffer_size;
  /* _reconfigure(), baseband.hh:156-194 */
  if (s->oFs > 0) {
    s->sub_sample = (size_t)(s->Fs / s->oFs);
    if (s->sub_sample < 1) s->sub_sample = 1;
  }
  update_filter_kernel(s);
  s->nco_Fs = (double)s->Fs;          /* setSampleRate(_Fs) */
  update_lut_incr(s);
  s->out_bs = s->source_bs / s->sub_sample;
  if (s->source_bs % s->sub_sample) s->out_bs += 1;
  s->last_r = s->last_i = 0; s->lastd_r = s->lastd_i = 0;
  s->sample_count = 0; s->ring_offset = 0;
  /* `_Fs/_sub_sample` (baseband.hh:193) is int32_t / size_t => unsigned 64-bit integer
   * division, then converted to double. */
  s->out_rate = (double)((size_t)s->Fs / s->sub_sample);
  return 0;
}

/* applyFrequencyShift for Scalar=int16 (CSScalar = complex<int32_t>), freqshift.hh:58-74 */
static inline void nco_s16(orc_iqbb *s, int32_t *vr, int32_t *vi) {
  if (0 == s->lut_inc) return;
  size_t idx = s->lut_count >> 8;
  if (0 > s->freq_shift) idx = ORC_LUT_SIZE - idx - 1;
  int32_t lr = s->lut_r[idx], li = s->lut_i[idx];
  int32_t re = w32_sub(w32_mul(lr, *vr), w32_mul(li, *vi));
  int32_t im = w32_add(w32_mul(lr, *vi), w32_mul(li, *vr));
  *vr = asr32(re, 16); *vi = asr32(im, 16);
  s->lut_count += s->lut_inc;
  while (s->lut_count >= (ORC_LUT_SIZE << 8)) s->lut_count -= (ORC_LUT_SIZE << 8);
}

/* applyFrequencyShift for Scalar=int8: CSScalar = complex<int16_t>.  The int32 FIR result is
 * narrowed to int16 on the call, the product is formed in int (promotion) and narrowed to int16
 * (libstdc++ complex<_Tp>::operator*=), the shift operator (operators.hh:48-50) widens to int32,
 * shifts by 8 and the result is narrowed to int16 again on return. */
static inline void nco_s8(orc_iqbb *s, int32_t *vr, int32_t *vi) {
  int16_t yr = (int16_t)*vr, yi = (int16_t)*vi;
  if (0 == s->lut_inc) { *vr = yr; *vi = yi; return; }
  size_t idx = s->lut_count >> 8;
  if (0 > s->freq_shift) idx = ORC_LUT_SIZE - idx - 1;
  int16_t lr = (int16_t)s->lut_r[idx], li = (int16_t)s->lut_i[idx];
  int16_t re = (int16_t)w32_sub(w32_mul(lr, yr), w32_mul(li, yi));
  int16_t im = (int16_t)w32_add(w32_mul(lr, yi), w32_mul(li, yr));
  *vr = (int16_t)asr32((int32_t)re, 8); *vi = (int16_t)asr32((int32_t)im, 8);
  s->lut_count += s->lut_inc;
  while (s->lut_count >= (ORC_LUT_SIZE << 8)) s->lut_count -= (ORC_LUT_SIZE << 8);
}

/* libstdc++ complex<int32_t>::operator/= with divisor (ss, 0):
 *   r = re*ss + im*0;  n = ss*ss + 0*0;  im' = (im*ss - re*0)/n;  re' = r/n   (all wrap) */
static inline void cdiv_ss(int32_t *re, int32_t *im, int32_t ss) {
  int32_t r = w32_add(w32_mul(*re, ss), w32_mul(*im, 0));
  int32_t n = w32_add(w32_mul(ss, ss), 0);
  int32_t i = w32_sub(w32_mul(*im, ss), w32_mul(*re, 0));
  /* INT_MIN / -1 cannot occur: n = ss*ss is never -1 */
  *im = (n != 0) ? i / n : 0;
  *re = (n != 0) ? r / n : 0;
}

static size_t process_int(orc_iqbb *s, const void *in, size_t n, void *out) {
  const int16_t *in16 = (const int16_t *)in; int16_t *out16 = (int16_t *)out;
  const int8_t *in8 = (const int8_t *)in; int8_t *out8 = (int8_t *)out;
  size_t j = 0;
  const size_t L = s->order;
  for (size_t i = 0; i < n; i++, s->sample_count++) {
    if (s->scalar == ORC_S16) { s->ring_r[s->ring_offset] = in16[2 * i]; s->ring_i[s->ring_offset] = in16[2 * i + 1]; }
    else { s->ring_r[s->ring_offset] = in8[2 * i]; s->ring_i[s->ring_offset] = in8[2 * i + 1]; }
    /* _filter_ring(), baseband.hh:226-236 */
    int32_t fr = 0, fi = 0;
    size_t idx = s->ring_offset + 1;
    if (L == idx) idx = 0;
    for (size_t t = 0; t < L; t++, idx++) {
      if (L == idx) idx = 0;
      int32_t xr = s->ring_r[idx], xi = s->ring_i[idx];
      fr = w32_add(fr, w32_sub(w32_mul(s->kr[t], xr), w32_mul(s->ki[t], xi)));
      fi = w32_add(fi, w32_add(w32_mul(s->kr[t], xi), w32_mul(s->ki[t], xr)));
    }
    fr = asr32(fr, 14); fi = asr32(fi, 14);
    if (s->scalar == ORC_S16) nco_s16(s, &fr, &fi); else nco_s8(s, &fr, &fi);
    s->last_r = w32_add(s->last_r, fr); s->last_i = w32_add(s->last_i, fi);
    s->ring_offset++;
    if (L == s->ring_offset) s->ring_offset = 0;
    if (s->sub_sample == s->sample_count) {
      int32_t vr = s->last_r, vi = s->last_i;
      cdiv_ss(&vr, &vi, (int32_t)s->sub_sample);
      if (s->scalar == ORC_S16) { out16[2 * j] = (int16_t)vr; out16[2 * j + 1] = (int16_t)vi; }
      else { out8[2 * j] = (int8_t)vr; out8[2 * j + 1] = (int8_t)vi; }
      s->last_r = s->last_i = 0; s->sample_count = 0; j++;
    } else if (s->sub_sample == 1) {
      if (s->scalar == ORC_S16) { out16[2 * j] = (int16_t)s->last_r; out16[2 * j + 1] = (int16_t)s->last_i; }
      else { out8[2 * j] = (int8_t)s->last_r; out8[2 * j + 1] = (int8_t)s->last_i; }
      s->last_r = s->last_i = 0; s->sample_count = 0; j++;
    }
  }
  return j;
}

/* Float variant.  DEFINED here (the reference does not compile for float): same structure with
 * SScalar=double arithmetic, kernel alpha/norm, LUT exp(-2 pi i j/128), no shifts, same 15-bit
 * phase accumulator and window grid, division by ss; result rounded to float on store. */
static size_t process_f32(orc_iqbb *s, const float *in, size_t n, float *out) {
  size_t j = 0;
  const size_t L = s->order;
  for (size_t i = 0; i < n; i++, s->sample_count++) {
    s->ringd_r[s->ring_offset] = in[2 * i]; s->ringd_i[s->ring_offset] = in[2 * i + 1];
    double fr = 0, fi = 0;
    size_t idx = s->ring_offset + 1;
    if (L == idx) idx = 0;
    for (size_t t = 0; t < L; t++, idx++) {
      if (L == idx) idx = 0;
      double xr = s->ringd_r[idx], xi = s->ringd_i[idx];
      fr += s->kdr[t] * xr - s->kdi[t] * xi;
      fi += s->kdr[t] * xi + s->kdi[t] * xr;
    }
    if (0 != s->lut_inc) {
      size_t li = s->lut_count >> 8;
      if (0 > s->freq_shift) li = ORC_LUT_SIZE - li - 1;
      double lr = s->lutd_r[li], lim = s->lutd_i[li];
      double re = lr * fr - lim * fi, im = lr * fi + lim * fr;
      fr = re; fi = im;
      s->lut_count += s->lut_inc;
      while (s->lut_count >= (ORC_LUT_SIZE << 8)) s->lut_count -= (ORC_LUT_SIZE << 8);
    }
    s->lastd_r += fr; s->lastd_i += fi;
    s->ring_offset++;
    if (L == s->ring_offset) s->ring_offset = 0;
    if (s->sub_sample == s->sample_count) {
      out[2 * j] = (float)(s->lastd_r / (double)s->sub_sample);
      out[2 * j + 1] = (float)(s->lastd_i / (double)s->sub_sample);
      s->lastd_r = s->lastd_i = 0; s->sample_count = 0; j++;
    } else if (s->sub_sample == 1) {
      out[2 * j] = (float)s->lastd_r; out[2 * j + 1] = (float)s->lastd_i;
      s->lastd_r = s->lastd_i = 0; s->sample_count = 0; j++;
    }
  }
  return j;
}

size_t orc_iqbb_process(orc_iqbb *s, const void *in, size_t n, void *out) {
  if (s->scalar == ORC_F32) return process_f32(s, (const float *)in, n, (float *)out);
  return process_int(s, in, n, out);
}

/* ---- real-input BaseBand<int16_t> (src/baseband.hh:304-529) --------------------------------- */
void orc_rbb_init(orc_iqbb *s, double Fc, double Ff, double width, size_t order, size_t sub_sample) {
  memset(s, 0, sizeof(*s));
  s->scalar = ORC_S16;
  s->freq_shift = Fc;
  s->order = order < 1 ? 1 : order;
  if (s->order > ORC_MAX_ORDER) s->order = ORC_MAX_ORDER;
  s->sub_sample = sub_sample;
  s->kdr[0] = Ff; s->kdr[1] = width;            /* parked here: the design needs the un-truncated values */
  build_lut(s);
}

int orc_rbb_config(orc_iqbb *s, double Fs, size_t buffer_size) {
  if (Fs == 0 || buffer_size == 0) return 1;
  const double rbb_Ff = s->kdr[0], rbb_width = s->kdr[1];   /* doubles in BaseBand, not int32 like IQBaseBand's */
  s->nco_Fs = Fs;                                /* setSampleRate(src_cfg.sampleRate()): a double */
  update_lut_incr(s);
  /* _update_filter_kernel, baseband.hh:462-487 */
  {
    static double ar[ORC_MAX_ORDER], ai[ORC_MAX_ORDER];
    double w = (2 * M_PI * rbb_width) / (2 * Fs);
    double M = (double)(s->order) / 2;
    double norm = 0;
    for (size_t i = 0; i < s->order; i++) {
      double a0 = (s->order == (2 * i)) ? 1 : sin(w * (i - M)) / (w * (i - M));
      double complex e = cexp(CMPLX(0, (2 * M_PI * rbb_Ff * i) / Fs));
      double re = a0 * creal(e) - 0.0 * cimag(e), im = a0 * cimag(e) + 0.0 * creal(e);
      double win = (0.42 - 0.5 * cos((2 * M_PI * (i + 1)) / (s->order + 2)) + 0.08 * cos((4 * M_PI * (i + 1)) / (s->order + 2)));
      re *= win; im *= win;
      ar[i] = re; ai[i] = im;
      norm += hypot(re, im);
    }
    for (size_t i = 0; i < s->order; i++) {
      const int sh = s->scalar == ORC_S8 ? 8 : 16;                /* 1 << Traits<Scalar>::shift */
      s->kr[i] = (int32_t)(((double)(1 << sh) * ar[i]) / norm);
      s->ki[i] = (int32_t)(((double)(1 << sh) * ai[i]) / norm);
    }
  }
  s->out_bs = buffer_size / s->sub_sample + ((buffer_size % s->sub_sample) ? 1 : 0);
  s->out_rate = Fs / (double)s->sub_sample;
  s->last_r = s->last_i = 0; s->sample_count = 0; s->ring_offset = 0;
  return 0;
}

/* BaseBand<int8_t>: the same class with Scalar = int8_t, i.e. SScalar = int16_t and CSScalar = complex<int16_t>
 * (src/freqshift.hh:20-22).  Everything the int16 instantiation does in 32 bits happens in 16 here:
 *   - the kernel is complex<int16_t>(2^8 alpha / norm), the FIR sum `res += _kernel[i] * _ring[idx]` wraps at 16 bits
 *     on every step (complex<int16_t> * int16_t and += narrow back to int16_t), then >> 8 (baseband.hh:443-453);
 *   - the mixer is FreqShiftBase<int8_t>'s (the int16 LUT, 16-bit product wrap: nco_s8 above);
 *   - `_last` is complex<int16_t>: the window sum wraps at 16 bits;
 *   - out = _last / CSScalar(sub_sample) is libstdc++'s complex<int16_t>::operator/=: the REAL part goes through
 *     `const _Tp __r = re * z.re + im * z.im` (narrowed to int16 BEFORE the division) while the imaginary part is
 *     (im * z.re - re * z.im) / n in int; n = std::norm(z) = int16(ss * ss); then narrowed to complex<int8_t>.
 * scalar must be set to ORC_S8 between orc_rbb_init and orc_rbb_config (orc_rbb_init8 does). */
void orc_rbb_init8(orc_iqbb *s, double Fc, double Ff, double width, size_t order, size_t sub_sample) {
  orc_rbb_init(s, Fc, Ff, width, order, sub_sample);
  s->scalar = ORC_S8;
  build_lut(s);
}
size_t orc_rbb_process8(orc_iqbb *s, const int8_t *in, size_t n, int8_t *out) {
  size_t j = 0;
  const size_t L = s->order;
  for (size_t i = 0; i < n; i++) {
    s->ring_r[s->ring_offset] = in[i];
    int16_t fr = 0, fi = 0;
    size_t idx = s->ring_offset + 1;
    if (L == idx) idx = 0;
    for (size_t t = 0; t < L; t++, idx++) {
      if (L == idx) idx = 0;
      const int16_t x = (int16_t)s->ring_r[idx];
      const int16_t pr = (int16_t)((int)(int16_t)s->kr[t] * (int)x), pi = (int16_t)((int)(int16_t)s->ki[t] * (int)x);
      fr = (int16_t)(fr + pr); fi = (int16_t)(fi + pi);
    }
    int32_t vr = (int16_t)(fr >> 8), vi = (int16_t)(fi >> 8);     /* >> Traits<int8_t>::shift on complex<int16_t> */
    nco_s8(s, &vr, &vi);
    s->last_r = (int16_t)(s->last_r + vr); s->last_i = (int16_t)(s->last_i + vi);
    s->sample_count++;
    s->ring_offset++;
    if (L == s->ring_offset) s->ring_offset = 0;
    if (s->sub_sample == s->sample_count) {
      const int16_t zr = (int16_t)s->sub_sample;                   /* CSScalar(_sub_sample) */
      const int16_t nn = (int16_t)((int)zr * (int)zr);             /* std::norm, narrowed to _Tp */
      const int16_t r = (int16_t)((int)(int16_t)s->last_r * (int)zr);
      const int16_t re = (int16_t)((int)r / (int)nn);
      const int16_t im = (int16_t)(((int)(int16_t)s->last_i * (int)zr) / (int)nn);
      out[2 * j] = (int8_t)re; out[2 * j + 1] = (int8_t)im;
      s->last_r = s->last_i = 0; s->sample_count = 0; j++;
    }
  }
  return j;
}

/* FreqShiftBase::setFrequencyShift, src/freqshift.hh:62-65 */
void orc_rbb_set_frequency_shift(orc_iqbb *s, double Fc) { s->freq_shift = Fc; update_lut_incr(s); }

size_t orc_rbb_process(orc_iqbb *s, const int16_t *in, size_t n, int16_t *out) {
  size_t j = 0;
  const size_t L = s->order;
  for (size_t i = 0; i < n; i++) {
    s->ring_r[s->ring_offset] = in[i];
    int32_t fr = 0, fi = 0;
    size_t idx = s->ring_offset + 1;
    if (L == idx) idx = 0;
    for (size_t t = 0; t < L; t++, idx++) {
      if (L == idx) idx = 0;
      fr = w32_add(fr, w32_mul(s->kr[t], s->ring_r[idx]));      /* complex<int32> * int32 */
      fi = w32_add(fi, w32_mul(s->ki[t], s->ring_r[idx]));
    }
    fr = asr32(fr, 16); fi = asr32(fi, 16);                      /* >> Traits<int16_t>::shift */
    nco_s16(s, &fr, &fi);
    s->last_r = w32_add(s->last_r, fr); s->last_i = w32_add(s->last_i, fi);
    s->sample_count++;
    s->ring_offset++;
    if (L == s->ring_offset) s->ring_offset = 0;
    if (s->sub_sample == s->sample_count) {
      int32_t vr = s->last_r, vi = s->last_i;
      cdiv_ss(&vr, &vi, (int32_t)s->sub_sample);
      out[2 * j] = (int16_t)vr; out[2 * j + 1] = (int16_t)vi;
      s->last_r = s->last_i = 0; s->sample_count = 0; j++;
    }
  }
  return j;
}

/* ---- demodulators ---------------------------------------------------------------------------- */

/* src/math.hh:12-21 and 31-40 (identical bodies once the operands are promoted to int32) */
int16_t orc_fast_atan2_i32(int32_t a, int32_t b) {
  const int32_t pi4 = (1 << 12), pi34 = 3 * (1 << 12);
  int32_t aabs, angle;
  if ((0 == a) && (0 == b)) return 0;
  aabs = (a >= 0) ? a : -a;
  if (b >= 0) angle = pi4 - pi4 * (b - aabs) / (b + aabs);
  else angle = pi34 - pi4 * (b + aabs) / (aabs - b);
  return (int16_t)((a >= 0) ? angle : -angle);
}

/* src/demod.hh:242-254 */
void orc_fmdemod_s16(const int16_t *in, size_t n, int16_t *out, int16_t *last) {
  for (size_t i = 1; i < n; i++) {
    int16_t re = in[2 * i], im = in[2 * i + 1];   /* read before a possible in-place write */
    int16_t phi = (int16_t)(orc_fast_atan2_i32(re, im) / 2);
    out[i] = (int16_t)(*last - phi);
    *last = phi;
  }
}

void orc_fmdemod_s8(const int8_t *in, size_t n, int16_t *out, int16_t *last) {
  for (size_t i = 1; i < n; i++) {
    int8_t re = in[2 * i], im = in[2 * i + 1];
    int16_t phi = (int16_t)(orc_fast_atan2_i32(re, im) / 2);
    out[i] = (int16_t)(*last - phi);
    *last = phi;
  }
}

double orc_fast_atan2_f64(double a, double b) {
  const double pi4 = M_PI / 4, pi34 = 3 * M_PI / 4;
  double aabs, angle;
  if ((0 == a) && (0 == b)) return 0;
  aabs = (a >= 0) ? a : -a;
  if (b >= 0) angle = pi4 - pi4 * (b - aabs) / (b + aabs);
  else angle = pi34 - pi4 * (b + aabs) / (aabs - b);
  return (a >= 0) ? angle : -angle;
}

void orc_fmdemod_f32(const float *in, size_t n, float *out, double *last) {
  for (size_t i = 1; i < n; i++) {
    double phi = orc_fast_atan2_f64(in[2 * i], in[2 * i + 1]) / 2;
    out[i] = (float)(*last - phi);
    *last = phi;
  }
}

/* src/demod.hh:65-81: sqrt(int) -> double -> Scalar */
void orc_amdemod_s16(const int16_t *in, size_t n, int16_t *out) {
  for (size_t i = 0; i < n; i++) {
    int32_t re = in[2 * i], im = in[2 * i + 1];
    int32_t q = w32_add(w32_mul(re, re), w32_mul(im, im));
    out[i] = (int16_t)(int32_t)sqrt((double)q);
  }
}
void orc_amdemod_s8(const int8_t *in, size_t n, int8_t *out) {
  for (size_t i = 0; i < n; i++) {
    int32_t re = in[2 * i], im = in[2 * i + 1];
    out[i] = (int8_t)(int32_t)sqrt((double)(re * re + im * im));
  }
}
void orc_amdemod_f32(const float *in, size_t n, float *out) {
  for (size_t i = 0; i < n; i++) {
    float re = in[2 * i], im = in[2 * i + 1];
    out[i] = sqrtf(re * re + im * im);
  }
}

/* src/demod.hh:156-161 */
void orc_usbdemod_s16(const int16_t *in, size_t n, int16_t *out) {
  for (size_t i = 0; i < n; i++) {
    int32_t re = in[2 * i], im = in[2 * i + 1];
    out[i] = (int16_t)((re + im) / 2);
  }
}
void orc_usbdemod_s8(const int8_t *in, size_t n, int8_t *out) {
  for (size_t i = 0; i < n; i++) {
    int16_t re = in[2 * i], im = in[2 * i + 1];
    out[i] = (int8_t)((re + im) / 2);
  }
}
void orc_usbdemod_f32(const float *in, size_t n, float *out) {
  for (size_t i = 0; i < n; i++) out[i] = (in[2 * i] + in[2 * i + 1]) / 2;
}

/* ---- AutoCast (src/autocast.hh:187-204) and FMDeemph (src/demod.hh:271-362) ------------------- */
void orc_autocast_u8_s16(const uint8_t *in, size_t n, int16_t *out) {
  const int8_t *v = (const int8_t *)in;           /* the reference reads uint8 data through int8_t* */
  for (size_t i = 0; i < n; i++) out[i] = (int16_t)(((int)(int16_t)v[i] - 127) << 8);
}
void orc_autocast_s8_s16(const int8_t *in, size_t n, int16_t *out) {
  for (size_t i = 0; i < n; i++) out[i] = (int16_t)((int)(int16_t)in[i] << 8);
}
int orc_fmdeemph_alpha(double Fs) { return (int)round(1.0 / ((1.0 - exp(-1.0 / (Fs * 75e-6))))); }
void orc_fmdeemph_s16(const int16_t *in, size_t n, int16_t *out, int alpha, int16_t *avg) {
  for (size_t i = 0; i < n; i++) {
    int16_t diff = (int16_t)(in[i] - *avg);       /* Scalar diff = in[i] - _avg */
    if (diff > 0) *avg = (int16_t)(*avg + (diff + alpha / 2) / alpha);
    else *avg = (int16_t)(*avg + (diff - alpha / 2) / alpha);
    out[i] = *avg;
  }
}

/* ---- FFT stand-in ---------------------------------------------------------------------------- */

static int is_pow2(size_t n) { return n && !(n & (n - 1)); }

void orc_fft_f64(const double *in, double *out, size_t n, int dir) {
  const double sgn = (dir > 0) ? -1.0 : 1.0;
  if (!is_pow2(n)) {
    double *tmp = (double *)malloc(2 * n * sizeof(double));
    for (size_t k = 0; k < n; k++) {
      double sr = 0, si = 0;
      for (size_t t = 0; t < n; t++) {
        double ang = sgn * 2 * M_PI * (double)((k * t) % n) / (double)n;
        double c = cos(ang), s = sin(ang);
        sr += in[2 * t] * c - in[2 * t + 1] * s;
        si += in[2 * t] * s + in[2 * t + 1] * c;
      }
      tmp[2 * k] = sr; tmp[2 * k + 1] = si;
    }
    memcpy(out, tmp, 2 * n * sizeof(double));
    free(tmp);
    return;
  }
  /* bit reversal copy */
  unsigned bits = 0; while (((size_t)1 << bits) < n) bits++;
  double *buf = (double *)malloc(2 * n * sizeof(double));
  for (size_t i = 0; i < n; i++) {
    size_t r = 0;
    for (unsigned b = 0; b < bits; b++) if (i & ((size_t)1 << b)) r |= (size_t)1 << (bits - 1 - b);
    buf[2 * r] = in[2 * i]; buf[2 * r + 1] = in[2 * i + 1];
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    size_t half = len >> 1;
    for (size_t k = 0; k < half; k++) {
      double ang = sgn * 2 * M_PI * (double)k / (double)len;
      double wr = cos(ang), wi = sin(ang);
      for (size_t s0 = 0; s0 < n; s0 += len) {
        size_t a = s0 + k, b = a + half;
        double tr = buf[2 * b] * wr - buf[2 * b + 1] * wi;
        double ti = buf[2 * b] * wi + buf[2 * b + 1] * wr;
        buf[2 * b] = buf[2 * a] - tr; buf[2 * b + 1] = buf[2 * a + 1] - ti;
        buf[2 * a] += tr; buf[2 * a + 1] += ti;
      }
    }
  }
  memcpy(out, buf, 2 * n * sizeof(double));
  free(buf);
}

void orc_fft_f32(const float *in, float *out, size_t n, int dir) {
  double *a = (double *)malloc(4 * n * sizeof(double));
  double *b = a + 2 * n;
  for (size_t i = 0; i < 2 * n; i++) a[i] = in[i];
  orc_fft_f64(a, b, n, dir);
  for (size_t i = 0; i < 2 * n; i++) out[i] = (float)b[i];
  free(a);
}

/* ---- overlap-add FFT filter ------------------------------------------------------------------ */

/* sinc_flt_kernel<float>, src/filternode.hh:17-28.  Note the precision choreography: the sinc is
 * evaluated in double and assigned to complex<float>; the phase factor is std::exp of a
 * complex<FLOAT> (argument rounded to float, cexpf); the window is computed in double and the
 * complex<float> is scaled by it via operator*=(const float&) (window rounded to float). */
void orc_filter_taps_f32(size_t block, double fmin_, double fmax_, double Fs, float *taps) {
  double fmin = fmin_ > -Fs / 2 ? fmin_ : -Fs / 2;     /* std::max(_fmin, -Fs/2) */
  double fmax = fmax_ < Fs / 2 ? fmax_ : Fs / 2;       /* std::min(_fmax, Fs/2) */
  double bw = fmax - fmin;
  double Fc = fmin + bw / 2;
  int N = (int)block;
  for (int i = 0; i < N; i++) {
    float vr, vi = 0.0f;
    if ((N / 2) == i) vr = (float)(M_PI * (bw / Fs));
    else vr = (float)(sin(M_PI * (bw / Fs) * (i - N / 2)) / (i - N / 2));
    float complex e = cexpf(CMPLXF(0.0f, (float)((2 * M_PI * Fc * i) / Fs)));
    float complex v = CMPLXF(vr, vi) * e;              /* complex<float> *= complex<float> */
    float win = (float)(0.42 - 0.5 * cos((2 * M_PI * i) / N) + 0.08 * cos((4 * M_PI * i) / N));
    taps[2 * i] = crealf(v) * win; taps[2 * i + 1] = cimagf(v) * win;
  }
}

/* FilterSource::_updateFilter, src/filternode.hh:186-203 (+ Buffer::norm2, buffer.hh:182-188,
 * Buffer::operator/=, buffer.hh:216-221: complex<float> /= complex<float>(norm)) */
void orc_filter_design_f32(size_t block, double fmin, double fmax, double Fs, float *kern) {
  size_t N = block;
  orc_filter_taps_f32(block, fmin, fmax, Fs, kern);
  for (size_t i = 0; i < 2 * N; i++) kern[2 * N + i] = 0.0f;
  orc_fft_f32(kern, kern, 2 * N, +1);
  double nrm2 = 0;
  for (size_t i = 0; i < 2 * N; i++) {
    /* std::real(std::conj(v)*v) in complex<float> arithmetic, accumulated in double */
    float re = kern[2 * i], im = kern[2 * i + 1];
    float p = re * re + im * im;
    nrm2 += p;
  }
  float nrm = (float)sqrt(nrm2);      /* `_kern /= _kern.norm2()`: double -> complex<float>(T) */
  for (size_t i = 0; i < 4 * N; i++) kern[i] = kern[i] / nrm;
}

void orc_filter_ola_block_f32(size_t block, const float *kern, const float *in, float *out, float *last) {
  size_t N = block, N2 = 2 * block;
  float *x = (float *)calloc(4 * N2, sizeof(float));
  float *X = x + 2 * N2;
  memcpy(x, in, 2 * N * sizeof(float));           /* FilterSink: first half data, second half 0 */
  orc_fft_f32(x, X, N2, +1);
  for (size_t i = 0; i < N2; i++) {               /* FilterSource: X[i]*K[i] in complex<float> */
    float ar = X[2 * i], ai = X[2 * i + 1], br = kern[2 * i], bi = kern[2 * i + 1];
    x[2 * i] = ar * br - ai * bi; x[2 * i + 1] = ar * bi + ai * br;
  }
  orc_fft_f32(x, X, N2, -1);
  float sc = (float)N2;
  for (size_t i = 0; i < N; i++) {
    out[2 * i] = last[2 * i] + X[2 * i] / sc;
    out[2 * i + 1] = last[2 * i + 1] + X[2 * i + 1] / sc;
    last[2 * i] = X[2 * (i + N)] / sc;
    last[2 * i + 1] = X[2 * (i + N) + 1] / sc;
  }
  free(x);
}
