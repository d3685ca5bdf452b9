"""GPU parity of the nodes either side of the hot path: AutoCast (stand-alone and
fused into IQBaseBand<int16_t>'s load) and FMDeemph -- bit-exact against the reference's goldens."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from libsdr_b200 import _lib, synth
from libsdr_b200.nodes import (IQBaseBand, RxChain, FMDeemph, autocast_cs16, Config, ConfigError, DEMOD_FM)
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cast_cu8", "cast_cs8"])
def test_autocast_golden(name):
    g = load_golden(name)
    np.testing.assert_array_equal(autocast_cs16(g["x"]), g["out"])
    import torch
    y = autocast_cs16(torch.from_numpy(g["x"]).cuda())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(y.cpu().numpy(), g["out"])


@pytest.mark.parametrize("dt,type_id", [(np.uint8, _lib.T_CU8), (np.int8, _lib.T_CS8)])
def test_autocast_fused_into_baseband(dt, type_id):
    """RTL-style 8-bit IQ straight into IQBaseBand<int16_t> == AutoCast node followed by IQBaseBand."""
    Fs, n, bs = 1e6, 6 * 16384, 16384
    t = np.arange(n) / Fs
    sig = 90 * np.exp(2j * np.pi * 103e3 * t) + 25 * np.exp(2j * np.pi * -200e3 * t + 1j)
    g = np.random.default_rng(5)
    raw = np.stack([sig.real, sig.imag], axis=1) + g.integers(-3, 4, size=(n, 2))
    x = (raw + 127).clip(0, 255).astype(np.uint8) if dt == np.uint8 else raw.clip(-128, 127).astype(np.int8)
    bb = IQBaseBand("s16", 100e3, 100e3, 12.5e3, 21, 1, 8000.0)
    bb.setInputType(type_id)
    bb.setCenterFrequency(100e3); bb.setFilterFrequency(100e3)
    out_cfg = bb.config(sample_rate=Fs, buffer_size=bs)
    assert out_cfg.type == _lib.T_CS16
    chain = RxChain(bb, DEMOD_FM)
    o = orc.IQBaseBand(orc.S16, 100e3, 100e3, 12.5e3, 21, 1, 8000.0)
    o.set_center_frequency(100e3); o.set_filter_frequency(100e3); o.config(Fs, bs)
    ofm = orc.FMDemod(orc.S16)
    xc = orc.autocast_cs16(x)
    # ragged calls so that the raw 8-bit history crosses call boundaries
    for s, e in [(0, bs), (bs, bs + 5), (bs + 5, 3 * bs), (3 * bs, n)]:
        yb, ya, _ = chain.process(x[s:e], e - s)
        ob = o.process(xc[s:e])
        np.testing.assert_array_equal(yb, ob)
        if ob.shape[0]:
            np.testing.assert_array_equal(ya, ofm.process(ob, inplace=True))
    with pytest.raises(ConfigError):
        bad = IQBaseBand("s16", 0.0, 0.0, 1e4, 15, 1, 0.0); bad.setInputType(_lib.T_CU8)
        bad.config(Config(_lib.T_CS16, 1e6, 1024, 1))            # now expects cu8


@pytest.mark.parametrize("name", golden_names("deemph_"))
def test_fmdeemph_golden(name):
    g = load_golden(name)
    d = FMDeemph(1)
    assert d.config(sample_rate=float(g["Fs"]), buffer_size=int(g["buffer_size"])).type == _lib.T_S16
    bs = int(g["buffer_size"])
    out = np.concatenate([d.process(g["x"][o:o + bs]) for o in range(0, g["x"].shape[0], bs)])
    np.testing.assert_array_equal(out, g["out"])


def test_fmdeemph_bank_of_streams():
    import torch
    streams, n = 300, 2000
    g = np.random.default_rng(1)
    x = g.integers(-20000, 20001, size=(streams, n)).astype(np.int16)
    d = FMDeemph(streams); d.config(sample_rate=48000.0, buffer_size=n)
    xd = torch.from_numpy(x).cuda()
    y1 = d.process(xd[:, :700].contiguous()); y2 = d.process(xd[:, 700:].contiguous())
    torch.cuda.synchronize()
    y = np.concatenate([y1.cpu().numpy(), y2.cpu().numpy()], axis=1)
    for s in (0, 1, 150, 299):
        o = orc.FMDeemph(48000.0)
        np.testing.assert_array_equal(y[s], o.process(x[s]))
    with pytest.raises(ConfigError):
        FMDeemph(1).config(Config(_lib.T_CS16, 48e3, 100, 1))
