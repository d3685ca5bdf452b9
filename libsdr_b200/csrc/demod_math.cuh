// demod_math.cuh -- per-sample device arithmetic of the demodulators and of the integer
// narrowing/division rules.  Semantics: src/math.hh:12-40, src/demod.hh:65-81,156-161,242-254,
// libstdc++ std::complex<int> operator/=.
#pragma once
#include <stdint.h>

namespace sdrg {

// fast_atan2<int16_t,int16_t> / <int8_t,int16_t>: int32 arithmetic, truncating division.
__device__ __forceinline__ int fast_atan2_int(int a, int b) {
  if ((a | b) == 0) return 0;
  const int aabs = a >= 0 ? a : -a;
  int angle;
  if (b >= 0) angle = 4096 - (4096 * (b - aabs)) / (b + aabs);
  else angle = 12288 - (4096 * (b + aabs)) / (aabs - b);
  return (int)(short)(a >= 0 ? angle : -angle);
}
// oScalar phi = fast_atan2(...)/2  (int division, then int16)
__device__ __forceinline__ int fm_phi_int(int re, int im) { return (int)(short)(fast_atan2_int(re, im) / 2); }

// Float FM (defined by this project): the same rational approximation in real
// arithmetic with pi/4 units; evaluated in double on the device (tiny, audio-rate work).
__device__ __forceinline__ double fm_phi_f64(double a, double b) {
  if (a == 0.0 && b == 0.0) return 0.0;
  const double pi4 = 0.78539816339744830962, pi34 = 2.35619449019234492885;
  const double aabs = fabs(a);
  double angle;
  if (b >= 0) angle = pi4 - pi4 * (b - aabs) / (b + aabs);
  else angle = pi34 - pi4 * (b + aabs) / (aabs - b);
  return (a >= 0 ? angle : -angle) / 2;
}

// AMDemod<int16_t>/<int8_t>: sqrt(int) in double, truncated, narrowed.
__device__ __forceinline__ int am_int(int re, int im) {
  const int q = (int)((unsigned)(re * re) + (unsigned)(im * im));
  return __double2int_rz(sqrt((double)q));
}
// USBDemod: (SScalar(re)+SScalar(im))/2, truncating
__device__ __forceinline__ int usb_int(int re, int im) { return (re + im) / 2; }

// complex<int32_t> /= complex<int32_t>(ss, 0) per component: wrap32(v*ss) / wrap32(ss*ss)
__device__ __forceinline__ int cdiv_component(int v, int ss) {
  const int num = (int)((unsigned)v * (unsigned)ss);
  const int den = (int)((unsigned)ss * (unsigned)ss);
  return den != 0 ? num / den : 0;
}

}  // namespace sdrg
