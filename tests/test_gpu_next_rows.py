"""GPU parity of the nodes either side of the hot path: AutoCast (stand-alone and
fused into IQBaseBand<int16_t>'s load) and FMDeemph -- bit-exact against the reference's goldens."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from libsdr_b200 import _lib, synth
from libsdr_b200.nodes import (IQBaseBand, BaseBand, RxChain, FMDeemph, autocast_cs16, Config, ConfigError, DEMOD_FM)
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


# ---- real-input BaseBand<int16_t> (src/baseband.hh:304-529) ---------------------------------------
@pytest.mark.parametrize("name", golden_names("rbb_"))
def test_real_baseband_golden(name):
    g = load_golden(name)
    bb = BaseBand(float(g["Fc"]), float(g["Ff"]), float(g["width"]), int(g["order"]), int(g["sub_sample"]))
    bs = int(g["buffer_size"])
    bb.config(sample_rate=float(g["Fs"]), buffer_size=bs)
    outs = [bb.process(g["x"][k:k + bs]) for k in range(0, g["x"].shape[0], bs)]
    np.testing.assert_array_equal(np.array([y.shape[0] for y in outs], dtype=np.uint32), g["counts"])
    np.testing.assert_array_equal(np.concatenate(outs), g["bb"])


@pytest.mark.parametrize("order,ss,Fc", [(15, 50, 300e3), (32, 1, -211e3), (40, 7, 0.0), (64, 300, 123456.7), (1, 3, 5e3)])
def test_real_baseband_ragged_vs_oracle(order, ss, Fc):
    """Arbitrary call boundaries (window carry + real history), full-scale input (int32 wrap in the FIR),
    tap counts inside and outside the fixed-tap kernel range; device tensors on the last cuts."""
    import torch
    Fs, n = 2.4e6, 200_000
    g = np.random.default_rng(order * 131 + ss)
    x = g.integers(-32768, 32768, size=n).astype(np.int16)
    bb = BaseBand(Fc, None, 90e3, order, ss); bb.config(sample_rate=Fs, buffer_size=65536)
    o = orc.BaseBand(Fc, Fc, 90e3, order, ss); o.config(Fs, 65536)
    cuts = [0, 1, 2, 2 + ss, 5000, 5001, 70000, 70000 + 3 * ss + 1, 150001, n]
    for k, (s, e) in enumerate(zip(cuts[:-1], cuts[1:])):
        if k >= 6:
            y = bb.process(torch.from_numpy(x[s:e]).cuda()); torch.cuda.synchronize(); y = y.cpu().numpy()
        else:
            y = bb.process(x[s:e])
        np.testing.assert_array_equal(y, o.process(x[s:e]))
    inf = bb.info()
    assert inf.samples_consumed == n and inf.outputs_produced == n // ss


def test_real_baseband_frequency_setter_and_fm_chain():
    """setFrequencyShift keeps the double (freqshift.hh:62-65) and restarts the phase; the complex output feeds
    the int16 FM demodulator like any IQBaseBand output."""
    from libsdr_b200.nodes import FMDemod
    Fs, n, ss = 192e3, 48000, 4
    t = np.arange(n) / Fs
    x = (12000 * np.cos(2 * np.pi * 30e3 * t + 3 * np.sin(2 * np.pi * 400 * t))).astype(np.int16)
    bb = BaseBand(10e3, 30e3, 16e3, 31, ss); bb.config(sample_rate=Fs, buffer_size=n)
    o = orc.BaseBand(10e3, 30e3, 16e3, 31, ss); o.config(Fs, n)
    np.testing.assert_array_equal(bb.process(x[:1000]), o.process(x[:1000]))
    bb.setFrequencyShift(30e3 + 0.625)
    o.set_frequency_shift(30e3 + 0.625)
    y, yo = bb.process(x[1000:]), o.process(x[1000:])
    np.testing.assert_array_equal(y, yo)
    fm, ofm = FMDemod("s16"), orc.FMDemod(orc.S16)
    fm.config(Config(_lib.T_CS16, Fs / ss, n // ss, 1))
    np.testing.assert_array_equal(fm.process(y)[1:], ofm.process(yo)[1:])


def test_autocast_whole_table():
    """Every cast of src/autocast.hh:30-69 on the device == the reference's bytes (golden) == the oracle; refused pairs
    raise ConfigError with the reference's message; host and device entry points."""
    import torch
    from conftest import load_golden
    from libsdr_b200.nodes import autocast
    g = load_golden("cast_table")
    x = g["x"]
    xd = torch.from_numpy(x).cuda()
    for k in g.files:
        if not k.startswith("y_"):
            continue
        _, i, o = k.split("_")
        np.testing.assert_array_equal(autocast(x, int(i), int(o)), g[k], err_msg=k)
        np.testing.assert_array_equal(autocast(xd, int(i), int(o)).cpu().numpy(), g[k], err_msg=k + " (device)")
    for i, o in g["refused"]:
        with pytest.raises(ConfigError, match="AutoCast: Can not cast"):
            autocast(x, int(i), int(o))
    big = np.random.default_rng(3).integers(0, 256, size=1 << 20).astype(np.uint8)
    for i, o in ((3, 10), (1, 8), (9, 8), (4, 2)):
        np.testing.assert_array_equal(autocast(big, i, o), orc.autocast(big, i, o))


@pytest.mark.parametrize("name", golden_names("rbb8_"))
def test_real_baseband_int8_golden(name):
    """BaseBand<int8_t> on the device == the reference's own outputs, per buffer; design (kernel, increment) identical."""
    g = load_golden(name)
    bs = int(g["buffer_size"])
    bb = BaseBand(float(g["Fc"]), float(g["Ff"]), float(g["width"]), int(g["order"]), int(g["sub_sample"]), scalar="s8")
    cfg = bb.config(sample_rate=float(g["Fs"]), buffer_size=bs)
    assert cfg.type == _lib.T_CS8
    k, _ = bb.design()
    np.testing.assert_array_equal(k, g["ref_kernel"])
    assert bb.info().lut_inc == int(g["ref_lut_inc"])
    x = g["x"]
    outs = [bb.process(x[o:o + bs]) for o in range(0, x.shape[0], bs)]
    np.testing.assert_array_equal(np.array([y.shape[0] for y in outs], dtype=np.uint32), g["counts"])
    np.testing.assert_array_equal(np.concatenate(outs), g["bb"])


def test_real_baseband_int8_random_and_errors():
    g = np.random.default_rng(81)
    for trial in range(10):
        Fs = float(g.choice([48e3, 1e6, 2.4e6])); order = int(g.integers(1, 41)); ss = int(g.choice([1, 2, 7, 16, 50, 100, 181, 200, 255, 300]))
        Fc = float(g.choice([0.0, 1.0, -1.0]) * g.uniform(0, 0.45) * Fs); Ff = float(g.uniform(-0.4, 0.4) * Fs); width = float(g.uniform(0.001, 0.4) * Fs)
        x = g.integers(-127, 128, size=30000).astype(np.int8)
        bb = BaseBand(Fc, Ff, width, order, ss, scalar="s8"); bb.config(sample_rate=Fs, buffer_size=4096)
        o = orc.BaseBand(Fc, Ff, width, order, ss, scalar=orc.S8); o.config(Fs, 4096)
        cuts = [0, 1, 4097, 20000, 30000]
        for s, e in zip(cuts[:-1], cuts[1:]):
            np.testing.assert_array_equal(bb.process(x[s:e]), o.process(x[s:e]), err_msg="trial %d" % trial)
    with pytest.raises(ConfigError):                    # int16(256 * 256) == 0: the reference would divide by zero
        BaseBand(1e3, 1e3, 500.0, 9, 256, scalar="s8").config(sample_rate=48e3, buffer_size=1024)
