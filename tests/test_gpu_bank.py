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


# ---- config 5: 2048 channels ---------------------------------------------------------------------------
C5_SPOT = (0, 1, 1023, 1024, 1025, 2047, 389, 1707)     # the edges, the centre (inc == 0) and two arbitrary ones


def c5_input(n, seed_tail=True):
    """C5's signal (amplitude 12 per carrier, noise +-8) with the carriers of the spot channels and their
    neighbours, followed (second half) by full-scale uniform noise so that the wrap regime of the window
    sums is exercised at 2048 channels as well."""
    cfg = synth.C5
    carriers = sorted({(k + d) % cfg["channels"] for k in C5_SPOT for d in (-1, 0, 1)} | set(range(0, 2048, 128)))
    x = synth.bank_input(n, cfg, carriers=carriers)
    if seed_tail:
        g = np.random.Generator(np.random.MT19937(0x5D120055))
        x[n // 2:] = g.integers(-32768, 32768, size=(n - n // 2, 2)).astype(np.int16)
    return x


def oracle_c5_channel(c, Fc, x, bs):
    cfg = synth.C5
    return oracle_channel("s16", Fc[c], Fc[c], cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"], cfg["Fs"], x, bs)


def check_c5_outputs(r, Fc, x, bs, channels=C5_SPOT):
    for c in channels:
        bb, f, a, u, first = oracle_c5_channel(c, Fc, x, bs)
        get = lambda k: (r[k][c].cpu().numpy() if hasattr(r[k], "cpu") else r[k][c])  # noqa: E731
        if "bb" in r:
            np.testing.assert_array_equal(get("bb"), bb, err_msg="bb ch %d" % c)
        np.testing.assert_array_equal(get("am"), a, err_msg="am ch %d" % c)
        mask = np.ones(f.shape[0], dtype=bool); mask[first] = False
        np.testing.assert_array_equal(get("fm")[mask], f[mask], err_msg="fm ch %d" % c)
        assert np.all(get("fm")[~mask] == 0)


def test_c5_shape_2048_channels():
    """BASELINE config 5 on ONE GPU: 2048 channels at C5's parameters (15 taps, 100 MS/s -> 48 kHz, amplitude 12),
    8 spot channels bit-exact against the oracle (base band, FM, AM), two buffers."""
    import torch
    cfg = synth.C5
    Fs, bs = cfg["Fs"], 1 << 17
    Fc = synth.bank_frequencies(cfg["channels"], Fs)
    x = c5_input(2 * bs)
    bank = ChannelBank("s16", Fc, None, cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"])
    bank.config(sample_rate=Fs, buffer_size=bs)
    r = bank.process(torch.from_numpy(x).cuda(), bs, want=("bb", "fm", "am"))
    torch.cuda.synchronize()
    assert r["fm"].shape == (2048, (2 * bs - 1) // 2083)
    check_c5_outputs(r, Fc, x, bs)


def _visible_devices():
    import torch
    return list(range(torch.cuda.device_count()))


@pytest.mark.parametrize("layout", ["same_device_x3", "all_devices"])
def test_sharded_bank_device_pointers(layout):
    """sdrg_bank_sharded_process_dev: input / outputs on devices[0], shards store their rows there.  On a
    one-GPU box the three shards share the device (the sharding, row placement and stream joins are the
    same code); with more GPUs visible every device takes a shard."""
    import torch
    from libsdr_b200.nodes import ShardedChannelBank
    devs = [0, 0, 0] if layout == "same_device_x3" else _visible_devices()
    if layout == "all_devices" and len(devs) < 2:
        pytest.skip("one GPU visible")
    cfg = synth.C5
    Fs, bs = cfg["Fs"], 1 << 16
    Fc = synth.bank_frequencies(cfg["channels"], Fs)
    x = c5_input(3 * bs)
    bank = ShardedChannelBank("s16", Fc, None, cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"], devices=devs)
    bank.config(sample_rate=Fs, buffer_size=bs)
    sh = bank.shards()
    assert [s[0] for s in sh] == devs and sum(s[2] for s in sh) == 2048 and sh[0][1] == 0
    if layout == "all_devices":
        assert all(s[3] for s in sh), "no peer access between the GPUs of this box: %r" % (sh,)
    xd = torch.from_numpy(x).to("cuda:%d" % devs[0])
    with torch.cuda.device(devs[0]):
        r1 = bank.process(xd[:2 * bs], bs, want=("fm", "am"))
        r1 = {k: v.clone() for k, v in r1.items()}
        r2 = bank.process(xd[2 * bs:], bs, want=("fm", "am"))       # carried state across calls, per shard
        torch.cuda.synchronize()
    r = {k: torch.cat([r1[k], r2[k]], dim=1) for k in r1}
    check_c5_outputs(r, Fc, x, bs)


def test_sharded_bank_host_pointers_equals_single_bank():
    from libsdr_b200.nodes import ShardedChannelBank
    devs = _visible_devices()
    devs = devs if len(devs) > 1 else [0, 0]
    Fs, bs = 2.4e6, 16384
    Fc = np.linspace(-1.1e6, 1.1e6, 37)
    x = synth.iq_int(4 * bs, Fs, [(8000, 103e3, 0.0), (5000, -98e3, 0.5), (3000, 340e3, 1.0)], 64, 9, np.int16)
    a = ChannelBank("s16", Fc, None, 25e3, 21, 300, 0.0); a.config(sample_rate=Fs, buffer_size=bs)
    b = ShardedChannelBank("s16", Fc, None, 25e3, 21, 300, 0.0, devices=devs); b.config(sample_rate=Fs, buffer_size=bs)
    for s, e in ((0, bs), (bs, 4 * bs)):
        ra, rb = a.process(x[s:e], bs), b.process(x[s:e], bs)
        for k in ra:
            np.testing.assert_array_equal(ra[k], rb[k], err_msg=k)


def test_sharded_bank_argument_errors():
    from libsdr_b200.nodes import ShardedChannelBank
    from libsdr_b200._lib import SDRError
    with pytest.raises(SDRError):
        ShardedChannelBank("s16", np.array([1e5, 2e5]), None, 25e3, 15, 300, 0.0, devices=[99])
    with pytest.raises(SDRError):
        ShardedChannelBank("s16", np.array([1e5]), None, 25e3, 15, 300, 0.0, devices=[0, 0])
