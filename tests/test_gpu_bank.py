"""GPU parity of the channel bank (configs C4/C5): every channel must equal, bit for bit, an
independent IQBaseBand<int16_t> -> {FM, AM, USB} (out of place) chain of the oracle."""
import numpy as np
import pytest

from libsdr_b200 import synth
from libsdr_b200.nodes import ChannelBank, ConfigError
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def oracle_channel(scalar, Fc, Ff, width, order, ss, oFs, Fs, x, bs):
    sc = orc.S16 if scalar == "s16" else orc.S8
    o = orc.IQBaseBand(sc, Fc, Ff, width, order, ss, oFs)
    o.config(Fs, bs)
    fm = orc.FMDemod(sc)
    bb, f, a, u, first = [], [], [], [], []
    off = 0
    for b in range(0, x.shape[0], bs):
        y = o.process(x[b:b + bs])
        bb.append(y)
        if y.shape[0]:
            f.append(fm.process(y, inplace=False)); first.append(off)
        a.append(orc.amdemod(y, sc)); u.append(orc.usbdemod(y, sc))
        off += y.shape[0]
    return np.concatenate(bb), np.concatenate(f), np.concatenate(a), np.concatenate(u), np.array(first, dtype=np.int64)


def check_bank(scalar, Fc, Ff, width, order, ss, oFs, Fs, x, bs, calls, channels_to_check=None):
    bank = ChannelBank(scalar, Fc, Ff, width, order, ss, oFs)
    bank.config(sample_rate=Fs, buffer_size=bs)
    outs = {k: [] for k in ("bb", "fm", "am", "usb")}
    for s, e in calls:
        r = bank.process(x[s:e], bs if (e - s) % bs == 0 else (e - s))
        for k in outs:
            outs[k].append(r[k].copy())
    got = {k: np.concatenate(v, axis=1) for k, v in outs.items()}
    for c in (range(len(Fc)) if channels_to_check is None else channels_to_check):
        bb, f, a, u, first = oracle_channel(scalar, Fc[c], Fc[c] if Ff is None else Ff[c], width, order, ss, oFs, Fs, x, bs)
        np.testing.assert_array_equal(got["bb"][c], bb, err_msg="bb ch %d" % c)
        np.testing.assert_array_equal(got["am"][c], a, err_msg="am ch %d" % c)
        np.testing.assert_array_equal(got["usb"][c], u, err_msg="usb ch %d" % c)
        mask = np.ones(f.shape[0], dtype=bool); mask[first] = False     # element 0 of each buffer: never written
        np.testing.assert_array_equal(got["fm"][c][mask], f[mask], err_msg="fm ch %d" % c)
        assert np.all(got["fm"][c][~mask] == 0)
    return bank


def test_small_bank_mixed_shifts():
    Fs, bs = 2.4e6, 16384
    Fc = np.array([100e3, -100e3, 0.0, 333e3, -777.5e3, 1.19e6, 50.25e3, -2.4e6])
    Ff = np.array([100e3, -100e3, 0.0, 300e3, -700e3, 1.0e6, 50e3, 0.0])
    x = synth.iq_int(6 * bs, Fs, [(8000, 103e3, 0.0), (5000, -98e3, 0.5), (3000, 340e3, 1.0), (2000, -770e3, 2.0)], 64, 3, np.int16)
    check_bank("s16", Fc, Ff, 25e3, 21, 300, 0.0, Fs, x, bs, [(0, 2 * bs), (2 * bs, 3 * bs), (3 * bs, 6 * bs)])


def test_bank_ragged_calls_and_wrap_regime():
    """ss=2083 like config 4, full-scale input (S*ss wraps 2^31), calls that cut windows anywhere."""
    Fs, bs = 100e6, 50000
    Fc = (np.arange(6) - 3) * (Fs / 256)
    x = synth.iq_int(4 * bs, Fs, [(20000, Fc[1] + 3e3, 0.0), (9000, Fc[4] - 2e3, 1.0)], 200, 4, np.int16)
    bank = ChannelBank("s16", Fc, None, 25e3, 15, 1, 48000.0)
    bank.config(sample_rate=Fs, buffer_size=bs)
    cuts = [0, 1, 2083, 2084, 50000, 123457, 200000]
    parts = [bank.process(x[s:e], e - s) for s, e in zip(cuts[:-1], cuts[1:])]
    bb = np.concatenate([p["bb"] for p in parts], axis=1)
    for c in range(6):
        o = orc.IQBaseBand(orc.S16, Fc[c], Fc[c], 25e3, 15, 1, 48000.0); o.config(Fs, bs)
        np.testing.assert_array_equal(bb[c], o.process(x))


def test_c4_shape_256_channels():
    cfg = synth.C4
    Fs, bs = cfg["Fs"], 1 << 18
    Fc = synth.bank_frequencies(cfg["channels"], Fs)
    x = synth.bank_input(2 * bs, dict(cfg, channels=16), chunk=1 << 16)      # 16 carriers are enough to excite the bank
    import torch
    xd = torch.from_numpy(x).cuda()
    bank = ChannelBank("s16", Fc, None, cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"])
    out_cfg = bank.config(sample_rate=Fs, buffer_size=bs)
    assert out_cfg.sample_rate == float(int(Fs) // 2083)
    r = bank.process(xd, bs, want=("bb", "fm", "am"))
    torch.cuda.synchronize()
    assert r["bb"].shape[0] == 256 and r["bb"].shape[1] == (2 * bs - 1) // 2083
    for c in (0, 1, 127, 128, 129, 200, 255):
        inf, k = bank.channel_info(c)
        o = orc.IQBaseBand(orc.S16, Fc[c], Fc[c], cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"]); o.config(Fs, bs)
        assert inf.lut_inc == o.lut_inc
        np.testing.assert_array_equal(k, o.kernel_i32())
        ob = np.concatenate([o.process(x[:bs]), o.process(x[bs:])])
        np.testing.assert_array_equal(r["bb"][c].cpu().numpy(), ob)
        np.testing.assert_array_equal(r["am"][c].cpu().numpy(), orc.amdemod(ob, orc.S16))


def test_int8_bank():
    Fs, bs = 2.4e6, 20000
    Fc = np.array([100e3, -300e3, 0.0])
    x = synth.iq_int(3 * bs, Fs, [(90, 101e3, 0.0), (30, -303e3, 1.0)], 5, 6, np.int8)
    check_bank("s8", Fc, None, 50e3, 17, 256, 0.0, Fs, x, bs, [(0, 3 * bs)])


def test_bank_rejects_small_subsampling():
    bank = ChannelBank("s16", np.array([1e5]), None, 25e3, 15, 50, 0.0)
    with pytest.raises(ConfigError):
        bank.config(sample_rate=2.4e6, buffer_size=4096)
