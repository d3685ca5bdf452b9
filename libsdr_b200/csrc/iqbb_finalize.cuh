// iqbb_finalize.cuh -- finalize (+ fused demodulation) of the completed windows of one call.
#pragma once
#include "iqbb_kernels.cuh"
#include "demod_math.cuh"

namespace sdrg {

template <int SCALAR> struct Fin;
template <> struct Fin<SDRG_T_S16> {
  typedef int2 Acc; typedef short2 Bb; typedef short Fm; typedef short Au; typedef int Last;
  static __device__ __forceinline__ int2 value(const Acc s, uint32_t ss) {
    if (ss == 1) return make_int2((int)(short)s.x, (int)(short)s.y);
    return make_int2((int)(short)cdiv_component(s.x, (int)ss), (int)(short)cdiv_component(s.y, (int)ss));
  }
  static __device__ __forceinline__ Bb store(int2 v) { return make_short2((short)v.x, (short)v.y); }
  static __device__ __forceinline__ Last phi(int2 v) { return fm_phi_int(v.x, v.y); }
  static __device__ __forceinline__ Fm fm(Last last, Last p) { return (short)(last - p); }
  static __device__ __forceinline__ Fm fm_first(int2 v) { return (short)v.x; }
  static __device__ __forceinline__ Au am(int2 v) { return (short)am_int(v.x, v.y); }
  static __device__ __forceinline__ Au usb(int2 v) { return (short)usb_int(v.x, v.y); }
};
template <> struct Fin<SDRG_T_S8> {
  typedef int2 Acc; typedef char2 Bb; typedef short Fm; typedef signed char Au; typedef int Last;
  static __device__ __forceinline__ int2 value(const Acc s, uint32_t ss) {
    if (ss == 1) return make_int2((int)(signed char)s.x, (int)(signed char)s.y);
    return make_int2((int)(signed char)cdiv_component(s.x, (int)ss), (int)(signed char)cdiv_component(s.y, (int)ss));
  }
  static __device__ __forceinline__ Bb store(int2 v) { return make_char2((signed char)v.x, (signed char)v.y); }
  static __device__ __forceinline__ Last phi(int2 v) { return fm_phi_int(v.x, v.y); }
  static __device__ __forceinline__ Fm fm(Last last, Last p) { return (short)(last - p); }
  static __device__ __forceinline__ Fm fm_first(int2 v) {   // int16 view of the two int8 bytes
    return (short)(((uint32_t)(uint8_t)v.x) | (((uint32_t)(uint8_t)v.y) << 8));
  }
  static __device__ __forceinline__ Au am(int2 v) { return (signed char)am_int(v.x, v.y); }
  static __device__ __forceinline__ Au usb(int2 v) { return (signed char)usb_int(v.x, v.y); }
};
template <> struct Fin<SDRG_T_F32> {
  typedef float2 Acc; typedef float2 Bb; typedef float Fm; typedef float Au; typedef double Last;
  static __device__ __forceinline__ float2 value(const Acc s, uint32_t ss) {
    if (ss == 1) return s;
    const float d = (float)ss;
    return make_float2(s.x / d, s.y / d);
  }
  static __device__ __forceinline__ Bb store(float2 v) { return v; }
  static __device__ __forceinline__ Last phi(float2 v) { return fm_phi_f64((double)v.x, (double)v.y); }
  static __device__ __forceinline__ Fm fm(Last last, Last p) { return (float)(last - p); }
  static __device__ __forceinline__ Fm fm_first(float2 v) { return v.x; }
  static __device__ __forceinline__ Au am(float2 v) { return sqrtf(v.x * v.x + v.y * v.y); }
  static __device__ __forceinline__ Au usb(float2 v) { return (v.x + v.y) / 2; }
};

// BaseBand<int8_t> (src/baseband.hh:304-529 with Scalar = int8_t): `_last` is complex<int16_t> and
// out = _last / complex<int16_t>(ss) is libstdc++'s complex<int16_t>::operator/=, whose real part is narrowed to
// int16 BEFORE the division (`const _Tp __r = ...`) while the imaginary part divides the int product; n = int16(ss ss).
__device__ __forceinline__ int2 fin_value_s8_real(const int2 s, uint32_t ss) {
  const int z = (int)(short)ss, nn = (int)(short)(z * z);
  if (nn == 0) return make_int2(0, 0);                          // refused at config(): the reference divides by zero
  const int sr = (int)(short)s.x, si = (int)(short)s.y;
  const int re = (int)(short)(((int)(short)(sr * z)) / nn);
  const int im = (int)(short)((si * z) / nn);
  return make_int2((int)(signed char)re, (int)(signed char)im);
}
template <int SCALAR>
__device__ __forceinline__ auto fin_value(const typename Fin<SCALAR>::Acc s, const IqbbFinalizeArgs &a) {
  if constexpr (SCALAR == SDRG_T_S8) { if (a.narrow16) return fin_value_s8_real(s, a.ss); }
  return Fin<SCALAR>::value(s, a.ss);
}

// is output j the first element of its segment (= of the buffer it is delivered with)?
// e = call-relative index of the sample that completes window j (< n <= 2^30, so 32-bit); the
// previous completion e - ss lies in an earlier buffer iff (e mod seg) < ss.
__device__ __forceinline__ bool seg_first(uint32_t j, const IqbbFinalizeArgs &a) {
  if (j == 0) return true;
  if (a.seg == 0) return false;
  const uint32_t e = a.e0 + j * a.ss;
  return (e % (uint32_t)a.seg) < a.ss;
}

// One CTA of 256 threads finalizes 256 consecutive windows.  FM needs the angle of the previous
// contributing sample: angles are exchanged through shared memory so that each one (a double
// precision division on the float path) is computed once; only a CTA's first thread, or a thread
// whose predecessor is a skipped buffer-first sample, recomputes from the accumulators.
template <int SCALAR>
__device__ __forceinline__ void iqbb_finalize_block(const IqbbFinalizeArgs &a, const uint32_t j0, const int tid,
                                                    typename Fin<SCALAR>::Last *sphi, unsigned char *sskip) {
  typedef Fin<SCALAR> F;
  const typename F::Acc *acc = (const typename F::Acc *)a.acc_cur;
  const uint32_t j = j0 + tid;
  const bool fm = a.demod == SDRG_DEMOD_FM;
  if (j == 0) {   // carry the open window and (folded float path) the tails already sent past it
    if (a.work_reset) *a.work_reset = 0u;
    ((typename F::Acc *)a.acc_next)[0] = acc[a.n_out];
    ((typename F::Acc *)a.acc_next)[1] = acc[a.n_out + 1];
    if (fm) {     // carried FM angle: the last sample of this call that contributes
      typename F::Last last = *(const typename F::Last *)a.fm_last_in;
      for (int64_t k = (int64_t)a.n_out - 1; k >= 0; --k) {
        if (!seg_first((uint32_t)k, a)) { last = F::phi(fin_value<SCALAR>(acc[k], a)); break; }
      }
      *(typename F::Last *)a.fm_last_out = last;
    }
  }
  const bool valid = j < a.n_out;
  auto v = fin_value<SCALAR>(valid ? acc[j] : typename F::Acc(), a);
  if (valid && a.bb_out) ((typename F::Bb *)a.bb_out)[j] = F::store(v);
  if (!a.audio_out) return;                      // uniform
  if (a.demod == SDRG_DEMOD_AM) { if (valid) ((typename F::Au *)a.audio_out)[j] = F::am(v); return; }
  if (a.demod == SDRG_DEMOD_USB) { if (valid) ((typename F::Au *)a.audio_out)[j] = F::usb(v); return; }
  if (!fm) return;
  const bool skip = !valid || seg_first(j, a);   // element 0 of a buffer is skipped (demod.hh:245)
  const typename F::Last p = skip ? typename F::Last(0) : F::phi(v);
  sphi[tid] = p; sskip[tid] = skip ? 1 : 0;
  __syncthreads();
  if (!valid) return;
  typename F::Fm *out = (typename F::Fm *)a.audio_out;
  if (skip) { if (a.in_place) out[j] = F::fm_first(v); return; }
  typename F::Last last;
  if (tid > 0 && !sskip[tid - 1]) last = sphi[tid - 1];
  else {
    int64_t k = (int64_t)j - 1;
    while (k >= 0 && seg_first((uint32_t)k, a)) --k;
    last = (k >= 0) ? F::phi(fin_value<SCALAR>(acc[k], a)) : *(const typename F::Last *)a.fm_last_in;
  }
  out[j] = F::fm(last, p);
}

}  // namespace sdrg
