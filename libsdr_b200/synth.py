"""Synthetic IQ workloads (host side, numpy) for the parity tests and bench.py.

These follow SURVEY.md section 8(d): a few complex tones evaluated in double at t = n/Fs with n the
global sample index, truncated toward zero to the integer type, plus seeded uniform noise.  The
bytes produced here are the single source of truth fed to both the CPU reference/oracle and the
GPU path.  (libsdr's own IQSigGen<int16_t> is not used: it emits amplitude-1 signals,
src/siggen.hh:98-107.)
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.MT19937(seed))


def tones(n, Fs, comps, start=0):
    """sum_k A_k exp(j(2 pi f_k t + p_k)), t=(start+i)/Fs, as complex128. comps = [(A, f, phase)]."""
    t = (np.arange(start, start + n, dtype=np.float64)) / Fs
    s = np.zeros(n, dtype=np.complex128)
    for (a, f, p) in comps:
        s += a * np.exp(1j * (2 * np.pi * f * t + p))
    return s


def iq_int(n, Fs, comps, noise, seed, dtype=np.int16, start=0):
    """Interleaved (n,2) integer IQ: trunc-toward-zero of the tone sum + uniform int noise, wrapped
    into dtype (so large amplitudes exercise the wrap paths)."""
    s = tones(n, Fs, comps, start)
    re = np.trunc(s.real).astype(np.int64)
    im = np.trunc(s.imag).astype(np.int64)
    if noise:
        g = _rng(seed)
        re += g.integers(-noise, noise + 1, size=n)
        im += g.integers(-noise, noise + 1, size=n)
    info = np.iinfo(dtype)
    out = np.stack([re, im], axis=1)
    out = np.clip(out, info.min, info.max)
    return out.astype(dtype)


def iq_f32(n, Fs, comps, noise, seed, start=0):
    s = tones(n, Fs, comps, start)
    g = _rng(seed)
    re = s.real + g.uniform(-noise, noise, size=n)
    im = s.imag + g.uniform(-noise, noise, size=n)
    return np.stack([re, im], axis=1).astype(np.float32)


# ---- the BASELINE.json configurations ------------------------------------------

C1 = dict(name="c1", scalar="s16", Fs=2.4e6, Fc=100e3, Ff=100e3, width=12.5e3, order=15, sub_sample=1,
          oFs=48000.0, buffer_size=65536, n_buffers=64)
C2 = dict(name="c2", scalar="f32", Fs=20e6, Fc=100e3, Ff=100e3, width=12.5e3, order=64, sub_sample=1,
          oFs=48000.0, buffer_size=1 << 20, n_buffers=64)
C3 = dict(name="c3", block=4096, Fs=20e6, fmin=100e3, fmax=300e3, buffer_size=1 << 20, n_buffers=16)
C4 = dict(name="c4", scalar="s16", Fs=100e6, width=25e3, order=15, sub_sample=1, oFs=48000.0,
          buffer_size=1 << 20, n_buffers=32, channels=256, amplitude=100, noise=8)
C5 = dict(name="c5", scalar="s16", Fs=100e6, width=25e3, order=15, sub_sample=1, oFs=48000.0,
          buffer_size=1 << 20, n_buffers=32, channels=2048, amplitude=12, noise=8)


def c1_input(n, start=0):
    comps = [(8192, 103e3, 0.0), (4096, 99e3, 0.5), (2730, 300e3, 1.0)]
    return iq_int(n, C1["Fs"], comps, 64, 0x5D120001, np.int16, start)


def c2_input(n, start=0):
    comps = [(0.5, 103e3, 0.0), (0.25, 99e3, 0.5), (0.1667, 300e3, 1.0)]
    return iq_f32(n, C2["Fs"], comps, 0.01, 0x5D120002, start)


def bank_frequencies(channels, Fs):
    """Carrier k sits at (k - C/2) Fs / C."""
    k = np.arange(channels)
    return (k - channels // 2) * (Fs / channels)


def bank_input(n, cfg, start=0, chunk=1 << 16, carriers=None):
    """C4/C5 wideband input: one FM-modulated carrier per channel (tone 1 kHz*(1+k mod 7),
    deviation 5 kHz, seeded phases), amplitude cfg['amplitude'], noise +-cfg['noise'], cs16.
    `carriers`: generate only these channel indices of the cfg['channels'] grid (a few dozen carriers
    excite a 2048-channel bank as well as all of them and cost 1/50 of the host time)."""
    Fs, C_ = cfg["Fs"], cfg["channels"]
    fk = bank_frequencies(C_, Fs)
    g = _rng(0x5D120004)
    ph = g.uniform(0, 2 * np.pi, size=C_)
    fm = 1e3 * (1 + (np.arange(C_) % 7))
    out = np.empty((n, 2), dtype=np.int16)
    for o in range(0, n, chunk):
        m = min(chunk, n - o)
        t = (np.arange(start + o, start + o + m, dtype=np.float64)) / Fs
        acc = np.zeros(m, dtype=np.complex128)
        for k in (range(C_) if carriers is None else carriers):
            acc += np.exp(1j * (2 * np.pi * fk[k] * t + ph[k] + (5e3 / fm[k]) * np.sin(2 * np.pi * fm[k] * t)))
        acc *= cfg["amplitude"]
        out[o:o + m, 0] = np.trunc(acc.real)
        out[o:o + m, 1] = np.trunc(acc.imag)
    gn = _rng(0x5D120005)
    out += gn.integers(-cfg["noise"], cfg["noise"] + 1, size=(n, 2)).astype(np.int16)
    return out
