"""Seeded random differential sweep: CUDA path vs oracle over random node parameters, amplitudes and
call boundaries.  Integer paths bit-exact (including the int32-wrap regime at full-scale input), float
within the stated tolerance."""
import numpy as np
import pytest

from libsdr_b200 import _lib
from libsdr_b200.nodes import IQBaseBand, BaseBand, RxChain, DEMOD_FM, DEMOD_AM, DEMOD_USB
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
FLOAT_TOL = 1e-5


def _cuts(g, n, k):
    c = np.unique(np.concatenate([[0, n], g.integers(1, n, size=k)]))
    return list(zip(c[:-1], c[1:]))


def _params(g):
    Fs = float(g.choice([48e3, 1e6, 2.4e6, 20e6, 100e6]))
    order = int(g.integers(1, 41))
    ss = int(g.choice([1, 2, 3, 7, 8, 31, 32, 50, 64, 65, 125, 256, 300, 2083]))
    Fc = float(g.choice([0.0, 1.0, -1.0]) * g.uniform(0, 0.45) * Fs)
    Ff = Fc if g.random() < 0.5 else float(g.uniform(-0.4, 0.4) * Fs)
    width = float(g.uniform(0.001, 0.4) * Fs)
    return Fs, Fc, Ff, width, order, ss


@pytest.mark.parametrize("seed", range(48))
@pytest.mark.parametrize("scalar", ["s16", "s8"])
def test_random_integer_chain(scalar, seed):
    g = np.random.default_rng(1000 * (scalar == "s8") + seed)
    Fs, Fc, Ff, width, order, ss = _params(g)
    n, bs = 40000, 8192
    dt = np.int16 if scalar == "s16" else np.int8
    full = g.random() < 0.4                               # full-scale noise: exercises the wrap regime
    amp = np.iinfo(dt).max if full else np.iinfo(dt).max // 8
    x = g.integers(-amp, amp + 1, size=(n, 2)).astype(dt)
    demod = [DEMOD_FM, DEMOD_AM, DEMOD_USB][seed % 3]
    bb = IQBaseBand(scalar, Fc, Ff, width, order, ss, 0.0); bb.config(sample_rate=Fs, buffer_size=bs)
    o = orc.IQBaseBand(orc.S16 if scalar == "s16" else orc.S8, Fc, Ff, width, order, ss, 0.0); o.config(Fs, bs)
    osc = orc.S16 if scalar == "s16" else orc.S8
    ofm = orc.FMDemod(osc)
    ch = RxChain(bb, demod)
    for s, e in _cuts(g, n, 5):
        yb, ya, _ = ch.process(x[s:e], e - s)
        ob = o.process(x[s:e])
        np.testing.assert_array_equal(yb, ob, err_msg="bb %s" % ((Fs, Fc, Ff, width, order, ss),))
        if ob.shape[0] == 0:
            continue
        if demod == DEMOD_FM:
            np.testing.assert_array_equal(ya[1:], ofm.process(ob, inplace=False)[1:])
        elif demod == DEMOD_AM:
            np.testing.assert_array_equal(ya, orc.amdemod(ob, osc))
        else:
            np.testing.assert_array_equal(ya, orc.usbdemod(ob, osc))


@pytest.mark.parametrize("seed", range(24))
def test_random_real_baseband(seed):
    g = np.random.default_rng(5000 + seed)
    Fs, Fc, Ff, width, order, ss = _params(g)
    n = 30000
    x = g.integers(-32768, 32768, size=n).astype(np.int16)
    bb = BaseBand(Fc, Ff, width, order, ss); bb.config(sample_rate=Fs, buffer_size=8192)
    o = orc.BaseBand(Fc, Ff, width, order, ss); o.config(Fs, 8192)
    for s, e in _cuts(g, n, 5):
        np.testing.assert_array_equal(bb.process(x[s:e]), o.process(x[s:e]), err_msg=str((Fs, Fc, Ff, width, order, ss)))


@pytest.mark.parametrize("seed", range(32))
def test_random_float_chain(seed):
    g = np.random.default_rng(9000 + seed)
    Fs, Fc, Ff, width, order, ss = _params(g)
    order = int(g.integers(1, 100))
    # a tone that the NCO moves exactly to DC keeps the window means away from zero (well-conditioned relative error)
    inc = int(g.integers(0, 12000))
    Fc = float(np.sign(Fc) or 1.0) * (inc + 0.5) * Fs / 32768 if inc else 0.0
    n = 60000
    t = np.arange(n)
    f_eff = np.sign(Fc) * inc * Fs / 32768
    x = (0.5 * np.exp(2j * np.pi * f_eff * t / Fs) + 0.05 * (g.standard_normal(n) + 1j * g.standard_normal(n))).astype(np.complex64)
    x = x.view(np.float32).reshape(-1, 2)
    bb = IQBaseBand("f32", Fc, Fc, max(width, 0.05 * Fs), order, ss, 0.0); bb.config(sample_rate=Fs, buffer_size=16384)
    o = orc.IQBaseBand(orc.F32, Fc, Fc, max(width, 0.05 * Fs), order, ss, 0.0); o.config(Fs, 16384)
    ys, os_ = [], []
    for s, e in _cuts(g, n, 4):
        ys.append(bb.process(x[s:e])); os_.append(o.process(x[s:e]))
        assert ys[-1].shape == os_[-1].shape
    y = np.concatenate(ys).astype(np.float64).view(np.complex128); ob = np.concatenate(os_).astype(np.float64).view(np.complex128)
    if ob.size:
        err = np.sqrt(np.mean(np.abs(y - ob) ** 2)) / max(np.sqrt(np.mean(np.abs(ob) ** 2)), 1e-30)
        assert err < FLOAT_TOL, (err, Fs, Fc, order, ss)
