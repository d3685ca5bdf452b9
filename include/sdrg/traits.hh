// sdrg/traits.hh -- scalar type table (src/traits.hh:20-236, src/traits.cc:6-34): for every
// buffer element type its real scalar, its compute "super scalar", the fixed-point shift and the
// full-scale value.  The shifts are part of the hot path's bit-exact contract.
#ifndef SDRG_TRAITS_HH
#define SDRG_TRAITS_HH

#include <complex>
#include <stdint.h>
#include <stddef.h>

#include "../sdrg.h"

namespace sdr {

template <class T> struct Traits;

#define SDRG_TRAIT(T, REAL, SUPER, SHIFT, SCALE, ID)                                              \
  template <> struct Traits<T> {                                                                  \
    typedef REAL Scalar; typedef std::complex<REAL> CScalar;                                      \
    typedef SUPER SScalar; typedef std::complex<SUPER> CSScalar;                                  \
    static constexpr size_t shift = SHIFT; static constexpr float scale = SCALE;                  \
    static constexpr int scalarId = ID;                                                           \
  };
SDRG_TRAIT(uint8_t, uint8_t, int16_t, 8, 127, SDRG_T_U8)
SDRG_TRAIT(int8_t, int8_t, int16_t, 8, 127, SDRG_T_S8)
SDRG_TRAIT(uint16_t, uint16_t, int32_t, 16, 32767, SDRG_T_U16)
SDRG_TRAIT(int16_t, int16_t, int32_t, 16, 32767, SDRG_T_S16)
SDRG_TRAIT(float, float, float, 0, 1, SDRG_T_F32)
SDRG_TRAIT(double, double, double, 0, 1, SDRG_T_F64)
SDRG_TRAIT(std::complex<uint8_t>, uint8_t, int16_t, 8, 127, SDRG_T_CU8)
SDRG_TRAIT(std::complex<int8_t>, int8_t, int16_t, 8, 127, SDRG_T_CS8)
SDRG_TRAIT(std::complex<uint16_t>, uint16_t, int32_t, 16, 32767, SDRG_T_CU16)
SDRG_TRAIT(std::complex<int16_t>, int16_t, int32_t, 16, 32767, SDRG_T_CS16)
SDRG_TRAIT(std::complex<float>, float, float, 0, 1, SDRG_T_CF32)
SDRG_TRAIT(std::complex<double>, double, double, 0, 1, SDRG_T_CF64)
#undef SDRG_TRAIT

}  // namespace sdr
#endif
