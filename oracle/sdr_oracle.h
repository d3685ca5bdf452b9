/* sdr_oracle.h -- CPU restatement of libsdr's receive-chain hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under libsdr_b200/ or include/ may include, link, load or
 * execute this file.  Only tests/, __graft_entry__.smoke() and bench.py (its cpu_baseline /
 * --impl reference legs, and the c5_check that compares the gathered multi-GPU bank output with
 * it OUTSIDE the timed region) use it, and only as the checker or the reported CPU baseline.
 *
 * Parity status: PINNED for the integer paths (IQBaseBand<int16_t>/<int8_t>, FMDemod, AMDemod,
 * USBDemod) against outputs of the reference itself, compiled from /root/reference/src by
 * oracle/Makefile into oracle/_ref/ref_harness (see tests/golden/gen_golden.py).
 * The float IQBaseBand/FMDemod have NO reference implementation (IQBaseBand<float> does not
 * compile, FMDemod<float> does not link): their semantics are DEFINED here -- "parity unpinned".
 * The overlap-add filter is pinned against the reference's FilterSink/FilterSource classes driven
 * with a double-precision stand-in for the absent FFTW plan.
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ORDER 1024
#define ORC_LUT_SIZE  128

enum { ORC_S8 = 0, ORC_S16 = 1, ORC_F32 = 2 };

/* IQBaseBand<Scalar> state (reference: src/baseband.hh:264-296, src/freqshift.hh:90-104). */
typedef struct {
  int      scalar;             /* ORC_S8 / ORC_S16 / ORC_F32 */
  /* constructor arguments */
  double   freq_shift;         /* FreqShiftBase::_freq_shift (double, NOT truncated by the ctor) */
  int32_t  Fc, Ff, Fs, width;  /* int32 members of IQBaseBand (baseband.hh:266-272) */
  size_t   order, sub_sample;
  double   oFs;
  /* derived at config() */
  double   nco_Fs;             /* FreqShiftBase::_Fs */
  size_t   lut_inc, lut_count;
  size_t   source_bs, out_bs;
  double   out_rate;
  int32_t  kr[ORC_MAX_ORDER], ki[ORC_MAX_ORDER];     /* integer kernel, 2^14 scaled */
  double   kdr[ORC_MAX_ORDER], kdi[ORC_MAX_ORDER];   /* float kernel (alpha/norm), double */
  int32_t  lut_r[ORC_LUT_SIZE], lut_i[ORC_LUT_SIZE]; /* int LUT, 2^shift scaled, trunc */
  double   lutd_r[ORC_LUT_SIZE], lutd_i[ORC_LUT_SIZE];
  /* running state */
  int32_t  ring_r[ORC_MAX_ORDER], ring_i[ORC_MAX_ORDER];
  double   ringd_r[ORC_MAX_ORDER], ringd_i[ORC_MAX_ORDER];
  size_t   ring_offset, sample_count;
  int32_t  last_r, last_i;
  double   lastd_r, lastd_i;
} orc_iqbb;

/* ctor: IQBaseBand(Fc, Ff, width, order, sub_sample, oFs)  (baseband.hh:35-57) */
void orc_iqbb_init(orc_iqbb *s, int scalar, double Fc, double Ff, double width,
                   size_t order, size_t sub_sample, double oFs);
/* setCenterFrequency / setFilterFrequency (baseband.hh:84-94) */
void orc_iqbb_set_center_frequency(orc_iqbb *s, double Fc);
void orc_iqbb_set_filter_frequency(orc_iqbb *s, double Ff);
/* config(): returns 0 ok; (baseband.hh:115-194) */
int  orc_iqbb_config(orc_iqbb *s, double sample_rate, size_t buffer_size);
/* _process(): consumes n complex samples, writes outputs, returns number written (baseband.hh:198-223).
 * in/out element type follows s->scalar: int8 pairs, int16 pairs, or float pairs. in==out allowed. */
size_t orc_iqbb_process(orc_iqbb *s, const void *in, size_t n, void *out);

/* FMDemod<int16,int16> / FMDemod<int8,int16> (demod.hh:242-254, math.hh:12-40).
 * out[0] is NOT written (reference quirk). last carried by the caller. */
int16_t orc_fast_atan2_i32(int32_t a, int32_t b);
void orc_fmdemod_s16(const int16_t *in_iq, size_t n, int16_t *out, int16_t *last);
void orc_fmdemod_s8(const int8_t *in_iq, size_t n, int16_t *out, int16_t *last);
/* float FM (DEFINED here, no reference): real-valued evaluation of math.hh:31-40 with pi4 = pi/4,
 * phi = atan2approx/2, out[i] = last - phi, i>=1; computed in double, stored as float. */
double orc_fast_atan2_f64(double a, double b);
void orc_fmdemod_f32(const float *in_iq, size_t n, float *out, double *last);

/* AMDemod (demod.hh:65-81) and USBDemod (demod.hh:156-161). */
void orc_amdemod_s16(const int16_t *in_iq, size_t n, int16_t *out);
void orc_amdemod_s8(const int8_t *in_iq, size_t n, int8_t *out);
void orc_amdemod_f32(const float *in_iq, size_t n, float *out);
void orc_usbdemod_s16(const int16_t *in_iq, size_t n, int16_t *out);
void orc_usbdemod_s8(const int8_t *in_iq, size_t n, int8_t *out);
void orc_usbdemod_f32(const float *in_iq, size_t n, float *out);

/* Real-input BaseBand<int16_t> (src/baseband.hh:304-529): complex band-pass FIR on a real stream
 * (kernel scaled 2^16, result >> 16), NCO, averaging decimator with plain ss-sample windows (the
 * sample counter is incremented BEFORE the test here, baseband.hh:425-436).  Reuses orc_iqbb's state:
 * ring_r holds the real history, kr/ki the kernel. */
void orc_rbb_init(orc_iqbb *s, double Fc, double Ff, double width, size_t order, size_t sub_sample);
int  orc_rbb_config(orc_iqbb *s, double sample_rate, size_t buffer_size);
void orc_rbb_set_frequency_shift(orc_iqbb *s, double Fc);
size_t orc_rbb_process(orc_iqbb *s, const int16_t *in, size_t n, int16_t *out_iq);
/* BaseBand<int8_t>: 16-bit arithmetic throughout (see sdr_oracle.c); sub_sample values whose square is 0 mod 2^16
 * divide by zero in the reference (SIGFPE) and must not be passed. */
void orc_rbb_init8(orc_iqbb *s, double Fc, double Ff, double width, size_t order, size_t sub_sample);
size_t orc_rbb_process8(orc_iqbb *s, const int8_t *in, size_t n, int8_t *out_iq);

/* AutoCast< std::complex<int16_t> > from complex 8-bit input (src/autocast.hh:187-204): n BYTES in,
 * n int16 out.  cu8: the bytes are read through an int8_t pointer (reference quirk), (v-127)<<8;
 * cs8: v<<8. */
void orc_autocast_u8_s16(const uint8_t *in, size_t n_bytes, int16_t *out);
void orc_autocast_s8_s16(const int8_t *in, size_t n_bytes, int16_t *out);
/* FMDeemph<int16_t> (src/demod.hh:271-362): alpha = round(1/(1-exp(-1/(Fs*75e-6)))); the running
 * average `avg` (int16) is carried by the caller and reset to 0 by config(). */
int  orc_fmdeemph_alpha(double sample_rate);
void orc_fmdeemph_s16(const int16_t *in, size_t n, int16_t *out, int alpha, int16_t *avg);

/* FFT (stand-in for the un-vendored FFTW3 behind src/fftplan_fftw3.hh:79-142): unnormalised DFT,
 * dir=+1 forward exp(-i..), dir=-1 backward exp(+i..). Power-of-two n: iterative radix-2 in double;
 * otherwise O(n^2) DFT in double. Interleaved re,im. */
void orc_fft_f64(const double *in, double *out, size_t n, int dir);
/* same but float in/out buffers, arithmetic in double (models FFTPlan<float>) */
void orc_fft_f32(const float *in, float *out, size_t n, int dir);

/* FilterSource::_updateFilter + sinc_flt_kernel<float> (filternode.hh:17-28,186-203):
 * writes the 2N-point normalised spectrum (float, interleaved) to kern. */
void orc_filter_design_f32(size_t block, double fmin, double fmax, double Fs, float *kern_2n);
/* time-domain taps only (h[i], i<N) exactly as sinc_flt_kernel<float> yields them */
void orc_filter_taps_f32(size_t block, double fmin, double fmax, double Fs, float *taps_n);
/* FilterSink::process + FilterSource::process (filternode.hh:81-88,164-181) for one block of N
 * samples; last_n is the carried overlap (N complex floats, zero-initialised by the caller). */
void orc_filter_ola_block_f32(size_t block, const float *kern_2n, const float *in_n,
                              float *out_n, float *last_n);

#ifdef __cplusplus
}
#endif
#endif
