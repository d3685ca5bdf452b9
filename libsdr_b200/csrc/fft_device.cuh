// fft_device.cuh -- register-level DFT building blocks shared by the FFT kernels (fft_kernels.cu, fft8k_kernels.cu,
// conv8k_kernels.cu).  Two flavours of the same functions: scalar FP32 (default) and, with SDRG_FFT_PACKED defined before
// the include, packed FP32 (sm_100 FADD2 / FMUL2 / FFMA2: one instruction per complex add, two per complex multiply).
#pragma once
#include <cuda_runtime.h>

namespace sdrg {
namespace {

#ifdef SDRG_FFT_PACKED

// Packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2 work on a register pair and take a half swap and per-half signs as operand
// modifiers, so one instruction does a complex add, and two do a complex multiply).
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {        // a b = a.x (b.x, b.y) + a.y (-b.y, b.x)
  return __ffma2_rn(make_float2(-b.y, b.x), make_float2(a.y, a.y), __fmul2_rn(b, make_float2(a.x, a.x)));
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// a * (c - i sn) forward, a * (c + i sn) inverse, c and sn compile-time constants
template <bool INV> __device__ __forceinline__ float2 mulw16(float2 a, const float c, const float sn) {
  const float2 r = rot90<INV>(a);                               // (a.y, -a.x) forward, (-a.y, a.x) inverse
  return __ffma2_rn(r, make_float2(sn, sn), __fmul2_rn(a, make_float2(c, c)));
}
template <bool INV> __device__ __forceinline__ void dft2(float2 &a, float2 &b) { const float2 t = a; a = caddf(t, b); b = csubf(t, b); }
// a + rot90(b), a - rot90(b): the rotation is the swap / sign modifier of one packed FMA each
template <bool INV> __device__ __forceinline__ void dft2_rot(float2 &a, float2 &b) {
  const float2 t = a, r = rot90<INV>(b), one = make_float2(1.f, 1.f);
  a = __ffma2_rn(r, one, t);
  b = __ffma2_rn(make_float2(-r.x, -r.y), one, t);
}
template <bool INV> __device__ __forceinline__ void dft4(float2 *v) {
  dft2<INV>(v[0], v[2]); dft2<INV>(v[1], v[3]);
  dft2<INV>(v[0], v[1]); dft2_rot<INV>(v[2], v[3]);
  const float2 t = v[1]; v[1] = v[2]; v[2] = t;       // bit reversal: outputs 0,2,1,3 -> natural
}
template <bool INV> __device__ __forceinline__ void dft8(float2 *v) {
  const float h = 0.70710678118654752440f;
  dft2<INV>(v[0], v[4]); dft2<INV>(v[1], v[5]); dft2<INV>(v[2], v[6]); dft2<INV>(v[3], v[7]);
  // twiddles w8^k on the odd half: 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2   (conjugated for the inverse)
  v[5] = mulw16<INV>(v[5], h, h);
  v[7] = mulw16<INV>(v[7], -h, h);
  dft2<INV>(v[0], v[2]); dft2<INV>(v[1], v[3]); dft2_rot<INV>(v[4], v[6]); dft2<INV>(v[5], v[7]);
  dft2<INV>(v[0], v[1]); dft2_rot<INV>(v[2], v[3]); dft2<INV>(v[4], v[5]); dft2_rot<INV>(v[6], v[7]);
  // outputs are in bit-reversed order 0,4,2,6,1,5,3,7
  float2 t;
  t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[3]; v[3] = v[6]; v[6] = t;
}

// 16-point DFT as 4 x 4 (Cooley-Tukey): DFT4 over r1 of v[4 r1 + r0], twiddle w16^(r0 q0), DFT4 over r0;
// result V[4 q1 + q0] in natural order.
template <bool INV> __device__ __forceinline__ void dft16(float2 *v) {
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  float2 a[4][4];                                   // a[r0][q0]
#pragma unroll
  for (int r0 = 0; r0 < 4; ++r0) {
    float2 t[4] = {v[r0], v[4 + r0], v[8 + r0], v[12 + r0]};
    dft4<INV>(t);
#pragma unroll
    for (int q0 = 0; q0 < 4; ++q0) a[r0][q0] = t[q0];
  }
  // w16^(r0 q0): exponents 1,2,3 / 2,4,6 / 3,6,9
  a[1][1] = mulw16<INV>(a[1][1], c1, s1); a[1][2] = mulw16<INV>(a[1][2], h, h);  a[1][3] = mulw16<INV>(a[1][3], s1, c1);
  a[2][1] = mulw16<INV>(a[2][1], h, h);   a[2][2] = rot90<INV>(a[2][2]);         a[2][3] = mulw16<INV>(a[2][3], -h, h);
  a[3][1] = mulw16<INV>(a[3][1], s1, c1); a[3][2] = mulw16<INV>(a[3][2], -h, h); a[3][3] = mulw16<INV>(a[3][3], -c1, -s1);
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) {
    float2 t[4] = {a[0][q0], a[1][q0], a[2][q0], a[3][q0]};
    dft4<INV>(t);
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) v[4 * q1 + q0] = t[q1];
  }
}


// a * w (forward) or a * conj(w) (inverse): tables hold the forward roots exp(-2 pi i k / n)
template <bool INV> __device__ __forceinline__ float2 cmulw(float2 a, float2 w) {
  // forward  a w       = w.x (a.x, a.y) + w.y (-a.y,  a.x)
  // inverse  a conj(w) = w.x (a.x, a.y) + w.y ( a.y, -a.x)
  const float2 r = INV ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
  return __ffma2_rn(r, make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}

#else

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <bool INV> __device__ __forceinline__ void dft2(float2 &a, float2 &b) { const float2 t = a; a = caddf(t, b); b = csubf(t, b); }
template <bool INV> __device__ __forceinline__ void dft4(float2 *v) {
  dft2<INV>(v[0], v[2]); dft2<INV>(v[1], v[3]);
  v[3] = rot90<INV>(v[3]);
  dft2<INV>(v[0], v[1]); dft2<INV>(v[2], v[3]);
  const float2 t = v[1]; v[1] = v[2]; v[2] = t;       // bit reversal: outputs 0,2,1,3 -> natural
}
template <bool INV> __device__ __forceinline__ void dft8(float2 *v) {
  const float h = 0.70710678118654752440f;
  dft2<INV>(v[0], v[4]); dft2<INV>(v[1], v[5]); dft2<INV>(v[2], v[6]); dft2<INV>(v[3], v[7]);
  // twiddles w8^k on the odd half: 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2   (conjugated for the inverse)
  v[5] = INV ? make_float2(h * (v[5].x - v[5].y), h * (v[5].x + v[5].y)) : make_float2(h * (v[5].x + v[5].y), h * (v[5].y - v[5].x));
  v[6] = rot90<INV>(v[6]);
  v[7] = INV ? make_float2(h * (-v[7].x - v[7].y), h * (v[7].x - v[7].y)) : make_float2(h * (v[7].y - v[7].x), h * (-v[7].x - v[7].y));
  dft2<INV>(v[0], v[2]); dft2<INV>(v[1], v[3]); dft2<INV>(v[4], v[6]); dft2<INV>(v[5], v[7]);
  v[3] = rot90<INV>(v[3]); v[7] = rot90<INV>(v[7]);
  dft2<INV>(v[0], v[1]); dft2<INV>(v[2], v[3]); dft2<INV>(v[4], v[5]); dft2<INV>(v[6], v[7]);
  // outputs are in bit-reversed order 0,4,2,6,1,5,3,7
  float2 t;
  t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[3]; v[3] = v[6]; v[6] = t;
}

// 16-point DFT as 4 x 4 (Cooley-Tukey): DFT4 over r1 of v[4 r1 + r0], twiddle w16^(r0 q0), DFT4 over r0;
// result V[4 q1 + q0] in natural order.
template <bool INV> __device__ __forceinline__ float2 mulw16(float2 a, const float c, const float sn) {
  // a * (c - i sn) forward, a * (c + i sn) inverse
  return INV ? make_float2(a.x * c - a.y * sn, a.y * c + a.x * sn) : make_float2(a.x * c + a.y * sn, a.y * c - a.x * sn);
}
template <bool INV> __device__ __forceinline__ void dft16(float2 *v) {
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  float2 a[4][4];                                   // a[r0][q0]
#pragma unroll
  for (int r0 = 0; r0 < 4; ++r0) {
    float2 t[4] = {v[r0], v[4 + r0], v[8 + r0], v[12 + r0]};
    dft4<INV>(t);
#pragma unroll
    for (int q0 = 0; q0 < 4; ++q0) a[r0][q0] = t[q0];
  }
  // w16^(r0 q0): exponents 1,2,3 / 2,4,6 / 3,6,9
  a[1][1] = mulw16<INV>(a[1][1], c1, s1); a[1][2] = mulw16<INV>(a[1][2], h, h);  a[1][3] = mulw16<INV>(a[1][3], s1, c1);
  a[2][1] = mulw16<INV>(a[2][1], h, h);   a[2][2] = rot90<INV>(a[2][2]);         a[2][3] = mulw16<INV>(a[2][3], -h, h);
  a[3][1] = mulw16<INV>(a[3][1], s1, c1); a[3][2] = mulw16<INV>(a[3][2], -h, h); a[3][3] = mulw16<INV>(a[3][3], -c1, -s1);
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) {
    float2 t[4] = {a[0][q0], a[1][q0], a[2][q0], a[3][q0]};
    dft4<INV>(t);
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) v[4 * q1 + q0] = t[q1];
  }
}


// a * w (forward) or a * conj(w) (inverse): tables hold the forward roots exp(-2 pi i k / n)
template <bool INV> __device__ __forceinline__ float2 cmulw(float2 a, float2 w) {
  return INV ? make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y) : make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

#endif

}  // namespace
}  // namespace sdrg
