// design.cc -- config-time (cold, host, double precision) coefficient design of IQBaseBand and
// FreqShiftBase.  The GPU kernels take these tables as data; they are never recomputed on the
// device because libm/device sin/cos would round differently and the integer paths must be
// bit-exact.  The arithmetic below follows, operation by operation, the expressions in
//   src/baseband.hh:239-262  (IQBaseBand::_update_filter_kernel)
//   src/freqshift.hh:26-36   (FreqShiftBase ctor: LUT)
//   src/freqshift.hh:78-87   (FreqShiftBase::_update_lut_incr)
// so that truncation to int32 lands on the same integers (checked against the reference's own
// tables: tests/test_abi_and_design.py on the golden vectors, tests/test_oracle_vs_live_reference.py on random
// parameter sets against the live reference harness).
#include "common.cuh"

#include <cmath>
#include <complex>

namespace sdrg {

static inline int trait_shift(int scalar) {      // Traits<Scalar>::shift, src/traits.cc:11-29
  return scalar == SDRG_T_S16 ? 16 : (scalar == SDRG_T_S8 ? 8 : 0);
}

void design_lut(IqbbDesign &d) {
  const double gain = double(1 << trait_shift(d.scalar));
  for (size_t j = 0; j < 128; ++j) {
    const std::complex<double> e = std::exp(std::complex<double>(0, -(2 * M_PI * j) / 128));
    const std::complex<double> v = gain * e;
    d.lutd_re[j] = v.real();
    d.lutd_im[j] = v.imag();
    if (d.scalar == SDRG_T_S8) {           // LUT element type is complex<int16_t> for int8 input
      d.lut_re[j] = int16_t(v.real());
      d.lut_im[j] = int16_t(v.imag());
    } else {
      d.lut_re[j] = int32_t(v.real());
      d.lut_im[j] = int32_t(v.imag());
    }
  }
}

void design_lut_increment(IqbbDesign &d, double nco_Fs) {
  d.negative = (0 > d.freq_shift);
  if (nco_Fs == 0) { d.lut_inc = 0; return; }
  const size_t lut_size = 128;
  d.lut_inc = size_t((lut_size * (1 << 8) * std::abs(d.freq_shift)) / nco_Fs);
}

void design_kernel(IqbbDesign &d) {
  const size_t L = d.order;
  std::complex<double> a[kMaxOrder];
  const double w = (M_PI * d.width) / (d.Fs);
  const double M = double(L) / 2.;
  double norm = 0;
  for (size_t i = 0; i < L; ++i) {
    if (L == 2 * i) a[i] = 4 * (w / M_PI);
    else a[i] = std::sin(w * (i - M)) / (w * (i - M));
    a[i] *= std::exp(std::complex<double>(0.0, (-2 * M_PI * d.Ff * i) / d.Fs));
    a[i] *= (0.42 - 0.5 * cos((2 * M_PI * i) / L) + 0.08 * cos((4 * M_PI * i) / L));
    norm += std::abs(a[i]);
  }
  for (size_t i = 0; i < L; ++i) {
    const std::complex<double> q = (double(1 << 14) * a[i]) / norm;
    d.k_re[i] = int32_t(q.real());
    d.k_im[i] = int32_t(q.imag());
    d.kd_re[i] = a[i].real() / norm;
    d.kd_im[i] = a[i].imag() / norm;
  }
}

// BaseBand<Scalar>::_update_filter_kernel (src/baseband.hh:462-487): centre tap 1, exp(+j..) shift,
// Blackman window over (i+1)/(order+2), gain 2^Traits<Scalar>::shift; Ff, width and Fs are doubles here.
void design_kernel_real(IqbbDesign &d, double Ff, double width, double Fs) {
  const size_t L = d.order;
  std::complex<double> a[kMaxOrder];
  const double w = (2 * M_PI * width) / (2 * Fs);
  const double M = double(L) / 2;
  double norm = 0;
  for (size_t i = 0; i < L; ++i) {
    if (L == 2 * i) a[i] = 1;
    else a[i] = std::sin(w * (i - M)) / (w * (i - M));
  }
  for (size_t i = 0; i < L; ++i) {
    a[i] = a[i] * std::exp(std::complex<double>(0, (2 * M_PI * Ff * i) / Fs));
    a[i] *= (0.42 - 0.5 * cos((2 * M_PI * (i + 1)) / (L + 2)) + 0.08 * cos((4 * M_PI * (i + 1)) / (L + 2)));
    norm += std::abs(a[i]);
  }
  for (size_t i = 0; i < L; ++i) {
    const std::complex<double> q = (double(1 << trait_shift(d.scalar)) * a[i]) / norm;
    d.k_re[i] = int32_t(q.real());
    d.k_im[i] = int32_t(q.imag());
    d.kd_re[i] = a[i].real() / norm;
    d.kd_im[i] = a[i].imag() / norm;
  }
}

const char *type_name(int type) {
  static const char *names[] = {"UNDEFINED", "uint8", "int8", "uint16", "int16", "float", "double",
                                "complex uint8", "complex int8", "complex uint16", "complex int16",
                                "complex float", "complex double"};
  return (type >= 0 && type <= SDRG_T_CF64) ? names[type] : "unknown";
}

}  // namespace sdrg
