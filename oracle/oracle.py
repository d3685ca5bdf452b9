"""ctypes binding of oracle/sdr_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (libsdr_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
REF_HARNESS = os.path.join(_HERE, "_ref", "ref_harness")

S8, S16, F32 = 0, 1, 2
_NP = {S8: np.int8, S16: np.int16, F32: np.float32}
MAX_ORDER = 1024


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref/ref_harness."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.DEVNULL if quiet else None)


def _load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "sdr_oracle.c")):
        subprocess.run(["make", "-s", "-C", _HERE, "_build/liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return C.CDLL(_LIB)


_lib = _load()


class _IQBB(C.Structure):
    _fields_ = [
        ("scalar", C.c_int),
        ("freq_shift", C.c_double),
        ("Fc", C.c_int32), ("Ff", C.c_int32), ("Fs", C.c_int32), ("width", C.c_int32),
        ("order", C.c_size_t), ("sub_sample", C.c_size_t),
        ("oFs", C.c_double),
        ("nco_Fs", C.c_double),
        ("lut_inc", C.c_size_t), ("lut_count", C.c_size_t),
        ("source_bs", C.c_size_t), ("out_bs", C.c_size_t),
        ("out_rate", C.c_double),
        ("kr", C.c_int32 * MAX_ORDER), ("ki", C.c_int32 * MAX_ORDER),
        ("kdr", C.c_double * MAX_ORDER), ("kdi", C.c_double * MAX_ORDER),
        ("lut_r", C.c_int32 * 128), ("lut_i", C.c_int32 * 128),
        ("lutd_r", C.c_double * 128), ("lutd_i", C.c_double * 128),
        ("ring_r", C.c_int32 * MAX_ORDER), ("ring_i", C.c_int32 * MAX_ORDER),
        ("ringd_r", C.c_double * MAX_ORDER), ("ringd_i", C.c_double * MAX_ORDER),
        ("ring_offset", C.c_size_t), ("sample_count", C.c_size_t),
        ("last_r", C.c_int32), ("last_i", C.c_int32),
        ("lastd_r", C.c_double), ("lastd_i", C.c_double),
    ]


_lib.orc_iqbb_init.argtypes = [C.POINTER(_IQBB), C.c_int, C.c_double, C.c_double, C.c_double,
                               C.c_size_t, C.c_size_t, C.c_double]
_lib.orc_iqbb_set_center_frequency.argtypes = [C.POINTER(_IQBB), C.c_double]
_lib.orc_iqbb_set_filter_frequency.argtypes = [C.POINTER(_IQBB), C.c_double]
_lib.orc_iqbb_config.argtypes = [C.POINTER(_IQBB), C.c_double, C.c_size_t]
_lib.orc_iqbb_config.restype = C.c_int
_lib.orc_iqbb_process.argtypes = [C.POINTER(_IQBB), C.c_void_p, C.c_size_t, C.c_void_p]
_lib.orc_iqbb_process.restype = C.c_size_t
_lib.orc_fast_atan2_i32.argtypes = [C.c_int32, C.c_int32]
_lib.orc_fast_atan2_i32.restype = C.c_int16
_lib.orc_fast_atan2_f64.argtypes = [C.c_double, C.c_double]
_lib.orc_fast_atan2_f64.restype = C.c_double
for _n in ("orc_fmdemod_s16", "orc_fmdemod_s8", "orc_fmdemod_f32"):
    getattr(_lib, _n).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
for _n in ("orc_amdemod_s16", "orc_amdemod_s8", "orc_amdemod_f32",
           "orc_usbdemod_s16", "orc_usbdemod_s8", "orc_usbdemod_f32"):
    getattr(_lib, _n).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
_lib.orc_fft_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
_lib.orc_fft_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
_lib.orc_filter_design_f32.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_void_p]
_lib.orc_filter_taps_f32.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_void_p]
_lib.orc_filter_ola_block_f32.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class IQBaseBand:
    """Oracle IQBaseBand<Scalar>. I/O arrays are interleaved (n, 2) of int8/int16/float32."""

    def __init__(self, scalar, Fc, Ff, width, order, sub_sample, oFs=0.0):
        self.s = _IQBB()
        self.scalar = scalar
        _lib.orc_iqbb_init(C.byref(self.s), scalar, Fc, Ff, width, order, sub_sample, oFs)

    def set_center_frequency(self, Fc):
        _lib.orc_iqbb_set_center_frequency(C.byref(self.s), Fc)

    def set_filter_frequency(self, Ff):
        _lib.orc_iqbb_set_filter_frequency(C.byref(self.s), Ff)

    def config(self, sample_rate, buffer_size):
        return _lib.orc_iqbb_config(C.byref(self.s), sample_rate, buffer_size)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=_NP[self.scalar]).reshape(-1, 2)
        ss = max(1, self.s.sub_sample)
        out = np.zeros((x.shape[0] // ss + 2, 2), dtype=_NP[self.scalar])
        n = _lib.orc_iqbb_process(C.byref(self.s), _p(x), x.shape[0], _p(out))
        return out[:n].copy()

    # design results
    @property
    def order(self): return self.s.order
    @property
    def sub_sample(self): return self.s.sub_sample
    @property
    def lut_inc(self): return self.s.lut_inc
    @property
    def neg(self): return self.s.freq_shift < 0
    @property
    def out_bs(self): return self.s.out_bs
    @property
    def out_rate(self): return self.s.out_rate

    def set_frequency_shift(self, Fc):
        _lib.orc_rbb_set_frequency_shift(C.byref(self.s), Fc)

    def kernel_i32(self):
        L = self.s.order
        return np.stack([np.array(self.s.kr[:L], dtype=np.int32), np.array(self.s.ki[:L], dtype=np.int32)], axis=1)

    def kernel_f64(self):
        L = self.s.order
        return np.array(self.s.kdr[:L]) + 1j * np.array(self.s.kdi[:L])

    def lut_i32(self):
        return np.stack([np.array(self.s.lut_r[:], dtype=np.int32), np.array(self.s.lut_i[:], dtype=np.int32)], axis=1)


_lib.orc_rbb_init.argtypes = [C.POINTER(_IQBB), C.c_double, C.c_double, C.c_double, C.c_size_t, C.c_size_t]
_lib.orc_rbb_config.argtypes = [C.POINTER(_IQBB), C.c_double, C.c_size_t]
_lib.orc_rbb_config.restype = C.c_int
_lib.orc_rbb_set_frequency_shift.argtypes = [C.POINTER(_IQBB), C.c_double]
_lib.orc_rbb_process.argtypes = [C.POINTER(_IQBB), C.c_void_p, C.c_size_t, C.c_void_p]
_lib.orc_rbb_process.restype = C.c_size_t
_lib.orc_rbb_init8.argtypes = [C.POINTER(_IQBB), C.c_double, C.c_double, C.c_double, C.c_size_t, C.c_size_t]
_lib.orc_rbb_process8.argtypes = [C.POINTER(_IQBB), C.c_void_p, C.c_size_t, C.c_void_p]
_lib.orc_rbb_process8.restype = C.c_size_t


class BaseBand:
    """Oracle real-input BaseBand<Scalar> (src/baseband.hh:304-529): real int16 in, (n,2) int16 out; with
    scalar=S8 the int8 instantiation (16-bit arithmetic throughout): real int8 in, (n,2) int8 out."""

    def __init__(self, Fc, Ff, width, order, sub_sample, scalar=S16):
        self.s = _IQBB()
        self.scalar = scalar
        (_lib.orc_rbb_init8 if scalar == S8 else _lib.orc_rbb_init)(C.byref(self.s), Fc, Ff, width, order, sub_sample)

    def config(self, sample_rate, buffer_size):
        return _lib.orc_rbb_config(C.byref(self.s), sample_rate, buffer_size)

    def process(self, x):
        dt = np.int8 if self.scalar == S8 else np.int16
        x = np.ascontiguousarray(x, dtype=dt).reshape(-1)
        out = np.zeros((x.shape[0] // max(1, self.s.sub_sample) + 2, 2), dtype=dt)
        fn = _lib.orc_rbb_process8 if self.scalar == S8 else _lib.orc_rbb_process
        n = fn(C.byref(self.s), _p(x), x.shape[0], _p(out))
        return out[:n].copy()

    def set_frequency_shift(self, Fc):
        _lib.orc_rbb_set_frequency_shift(C.byref(self.s), Fc)

    def kernel_i32(self):
        L = self.s.order
        return np.stack([np.array(self.s.kr[:L], dtype=np.int32), np.array(self.s.ki[:L], dtype=np.int32)], axis=1)

    @property
    def lut_inc(self): return self.s.lut_inc
    @property
    def out_rate(self): return self.s.out_rate
    @property
    def out_bs(self): return self.s.out_bs


class FMDemod:
    """Oracle FMDemod; `inplace` selects what index 0 of each buffer shows (demod.hh:229-254)."""

    def __init__(self, scalar):
        self.scalar = scalar
        self.last_i = np.zeros(1, dtype=np.int16)
        self.last_f = np.zeros(1, dtype=np.float64)

    def process(self, x, inplace=True):
        x = np.ascontiguousarray(x, dtype=_NP[self.scalar]).reshape(-1, 2)
        n = x.shape[0]
        if n == 0:
            return None
        if self.scalar == F32:
            out = np.zeros(n, dtype=np.float32)
            if inplace:
                out[0] = x[0, 0]
            _lib.orc_fmdemod_f32(_p(x), n, _p(out), _p(self.last_f))
            return out
        out = np.zeros(n, dtype=np.int16)
        if inplace:
            # the int16 view of the input's first two bytes
            out[0] = x.view(np.int16).reshape(-1)[0] if self.scalar == S16 else x.reshape(-1)[:2].view(np.int16)[0]
        fn = _lib.orc_fmdemod_s16 if self.scalar == S16 else _lib.orc_fmdemod_s8
        fn(_p(x), n, _p(out), _p(self.last_i))
        return out


def amdemod(x, scalar):
    x = np.ascontiguousarray(x, dtype=_NP[scalar]).reshape(-1, 2)
    out = np.zeros(x.shape[0], dtype=_NP[scalar])
    {S8: _lib.orc_amdemod_s8, S16: _lib.orc_amdemod_s16, F32: _lib.orc_amdemod_f32}[scalar](_p(x), x.shape[0], _p(out))
    return out


def usbdemod(x, scalar):
    x = np.ascontiguousarray(x, dtype=_NP[scalar]).reshape(-1, 2)
    out = np.zeros(x.shape[0], dtype=_NP[scalar])
    {S8: _lib.orc_usbdemod_s8, S16: _lib.orc_usbdemod_s16, F32: _lib.orc_usbdemod_f32}[scalar](_p(x), x.shape[0], _p(out))
    return out


_lib.orc_autocast_u8_s16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
_lib.orc_autocast_s8_s16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
_lib.orc_fmdeemph_alpha.argtypes = [C.c_double]
_lib.orc_fmdeemph_alpha.restype = C.c_int
_lib.orc_fmdeemph_s16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]


def autocast_cs16(x):
    """AutoCast<complex<int16>> of complex uint8 / int8 input (n,2) -> (n,2) int16."""
    x = np.ascontiguousarray(x)
    out = np.zeros(x.shape, dtype=np.int16)
    fn = _lib.orc_autocast_u8_s16 if x.dtype == np.uint8 else _lib.orc_autocast_s8_s16
    fn(_p(x), x.size, _p(out))
    return out


# ---- the whole AutoCast table (src/autocast.hh:30-69, cast functions :120-262) ---------------------------------
# Config::Type ids (src/node.hh:39-53)
T_U8, T_S8, T_U16, T_S16, T_CU8, T_CS8, T_CU16, T_CS16 = 1, 2, 3, 4, 7, 8, 9, 10


def _w(a, dt):
    """two's-complement narrowing like the reference's implicit conversions"""
    return np.asarray(a).astype(np.int64).astype({8: np.int8, 16: np.int16}[dt])


def _cplx(re):
    out = np.zeros((re.shape[0], 2), dtype=re.dtype); out[:, 0] = re
    return out.reshape(-1)


_CASTS = {
    # AutoCast<int8_t>
    (T_U8, T_S8): lambda b: _w(b.view(np.uint8).astype(np.int64) - 127, 8),                            # _uint8_int8, :137-144
    (T_U16, T_S8): lambda b: _w((b.view(np.uint16).astype(np.int64) >> 8) - 127, 8),                   # _uint16_int8
    (T_S16, T_S8): lambda b: _w(b.view(np.int16).astype(np.int64) >> 8, 8),                            # _int16_int8
    # AutoCast<complex<int8_t>>
    (T_U8, T_CS8): lambda b: _cplx(_w(b.view(np.uint8).astype(np.int64) - 127, 8)),                    # _uint8_cint8
    (T_S8, T_CS8): lambda b: _cplx(b.view(np.int8).copy()),                                            # _int8_cint8
    (T_U16, T_CS8): lambda b: _cplx(_w((b.view(np.int16).astype(np.int64) >> 8) - 65535, 8)),          # _uint16_cint8: read as int16, - ((2<<15)-1)
    (T_S16, T_CS8): lambda b: _cplx(_w(b.view(np.int16).astype(np.int64) >> 8, 8)),                    # _int16_cint8
    # AutoCast<int16_t>
    (T_U8, T_S16): lambda b: _w((b.view(np.int8).astype(np.int64) - 127) << 8, 16),                    # _uint8_int16: read through int8_t*
    (T_S8, T_S16): lambda b: _w(b.view(np.int8).astype(np.int64) << 8, 16),                            # _int8_int16
    (T_U16, T_S16): lambda b: _w(b.view(np.uint16).astype(np.int64) - 65535, 16),                      # _uint16_int16: - ((2<<15)-1)
    # AutoCast<complex<int16_t>>
    (T_U8, T_CS16): lambda b: _cplx(_w((b.view(np.uint8).astype(np.int64) - 127) << 8, 16)),           # _uint8_cint16 (uint8 here)
    (T_S8, T_CS16): lambda b: _cplx(_w(b.view(np.int8).astype(np.int64) * 256, 16)),                   # _int8_cint16
    (T_U16, T_CS16): lambda b: _cplx(_w(b.view(np.uint16).astype(np.int64) - 32768, 16)),              # _uint16_cint16: - (1<<15)
    (T_S16, T_CS16): lambda b: _cplx(b.view(np.int16).copy()),                                         # _int16_cint16
}
# complex sources reuse the scalar casts on the interleaved components (autocast.hh:44-68)
_CASTS[(T_CU8, T_CS8)] = _CASTS[(T_U8, T_S8)]
_CASTS[(T_CU16, T_CS8)] = _CASTS[(T_U16, T_S8)]
_CASTS[(T_CS16, T_CS8)] = _CASTS[(T_S16, T_S8)]
_CASTS[(T_CU8, T_CS16)] = _CASTS[(T_U8, T_S16)]
_CASTS[(T_CS8, T_CS16)] = _CASTS[(T_S8, T_S16)]
_CASTS[(T_CU16, T_CS16)] = _CASTS[(T_U16, T_S16)]
for _t in (T_S8, T_CS8, T_S16, T_CS16):
    _CASTS[(_t, _t)] = lambda b: b.copy()                                                              # _identity


def autocast_supported(in_type, out_type):
    return (int(in_type), int(out_type)) in _CASTS


def autocast(raw_bytes, in_type, out_type):
    """AutoCast<out_type> applied to a byte string holding elements of in_type; returns the output BYTES (uint8)."""
    b = np.ascontiguousarray(raw_bytes).view(np.uint8).reshape(-1)
    return np.ascontiguousarray(_CASTS[(int(in_type), int(out_type))](b)).view(np.uint8).reshape(-1)


class FMDeemph:
    def __init__(self, sample_rate):
        self.alpha = int(_lib.orc_fmdeemph_alpha(float(sample_rate)))
        self.avg = np.zeros(1, dtype=np.int16)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.zeros_like(x)
        _lib.orc_fmdeemph_s16(_p(x), x.size, _p(out), self.alpha, _p(self.avg))
        return out


def fast_atan2_i32(a, b):
    return int(_lib.orc_fast_atan2_i32(int(a), int(b)))


def fft_f64(x, direction=+1):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    _lib.orc_fft_f64(_p(x), _p(out), x.shape[0], direction)
    return out


def fft_f32(x, direction=+1):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    _lib.orc_fft_f32(_p(x), _p(out), x.shape[0], direction)
    return out


def filter_taps(block, fmin, fmax, Fs):
    taps = np.zeros(block, dtype=np.complex64)
    _lib.orc_filter_taps_f32(block, fmin, fmax, Fs, _p(taps))
    return taps


def filter_design(block, fmin, fmax, Fs):
    kern = np.zeros(2 * block, dtype=np.complex64)
    _lib.orc_filter_design_f32(block, fmin, fmax, Fs, _p(kern))
    return kern


class FilterOLA:
    """Oracle FilterSink<float> + FilterSource<float> (filternode.hh:81-88,164-181)."""

    def __init__(self, block, fmin, fmax, Fs):
        if fmax < fmin:
            fmin, fmax = fmax, fmin
        self.block = block
        self.kern = filter_design(block, fmin, fmax, Fs)
        self.last = np.zeros(block, dtype=np.complex64)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        assert x.shape[0] % self.block == 0
        out = np.empty_like(x)
        for b in range(x.shape[0] // self.block):
            sl = slice(b * self.block, (b + 1) * self.block)
            xin = np.ascontiguousarray(x[sl]); o = np.empty(self.block, dtype=np.complex64)
            _lib.orc_filter_ola_block_f32(self.block, _p(self.kern), _p(xin), _p(o), _p(self.last))
            out[sl] = o
        return out


def filter_timedomain_f64(taps_c64, x, history=None):
    """Independent time-domain oracle: y[n] = sum_j (h[j]/nrm) x[n-j],
    nrm = ||FFT_2N(h)||_2 = sqrt(2N) ||h||_2, in double."""
    h = taps_c64.astype(np.complex128)
    N = h.shape[0]
    nrm = np.sqrt(2 * N) * np.sqrt(np.sum(np.abs(h) ** 2))
    xx = x.astype(np.complex128)
    return np.convolve(xx, h / nrm)[: xx.shape[0]]


def run_ref(args):
    """Run oracle/_ref/ref_harness with the given argv; returns stdout."""
    if not os.path.exists(REF_HARNESS):
        raise FileNotFoundError(REF_HARNESS)
    return subprocess.run([REF_HARNESS] + [str(a) for a in args], check=True, capture_output=True, text=True).stdout
