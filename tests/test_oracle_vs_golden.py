"""Pins oracle/sdr_oracle.c against outputs of the reference itself (tests/golden/*.npz, produced
by tests/golden/gen_golden.py from the classes in /root/reference/src).  CPU only."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_rms
from oracle import oracle as orc

SC = {"s16": orc.S16, "s8": orc.S8}


def make_bb(g):
    bb = orc.IQBaseBand(SC[str(g["scalar"])], float(g["Fc"]), float(g["Ff"]), float(g["width"]),
                        int(g["order"]), int(g["sub_sample_arg"]), float(g["oFs"]))
    if int(g["setcf"]):
        bb.set_center_frequency(float(g["Fc"]))
        bb.set_filter_frequency(float(g["Ff"]))
    bb.config(float(g["Fs"]), int(g["buffer_size"]))
    return bb


@pytest.mark.parametrize("name", golden_names("bb_"))
def test_design_matches_reference(name):
    g = load_golden(name)
    bb = make_bb(g)
    assert bb.order == int(g["ref_order"])
    assert bb.sub_sample == int(g["ref_sub_sample"])
    assert bb.lut_inc == int(g["ref_lut_inc"])
    assert int(bb.neg) == int(g["ref_neg"])
    np.testing.assert_array_equal(bb.kernel_i32(), g["ref_kernel"])
    np.testing.assert_array_equal(bb.lut_i32(), g["ref_lut"])


@pytest.mark.parametrize("name", golden_names("bb_"))
def test_baseband_and_demods_bit_exact(name):
    g = load_golden(name)
    sc = SC[str(g["scalar"])]
    bb = make_bb(g)
    fm = orc.FMDemod(sc)
    x, bs = g["x"], int(g["buffer_size"])
    bbs, fms, ams, usbs, counts = [], [], [], [], []
    for off in range(0, x.shape[0], bs):
        y = bb.process(x[off:off + bs])
        counts.append(y.shape[0])
        bbs.append(y)
        if y.shape[0]:
            fms.append(fm.process(y, inplace=True))
        ams.append(orc.amdemod(y, sc)); usbs.append(orc.usbdemod(y, sc))
    np.testing.assert_array_equal(np.array(counts, dtype=np.uint32), g["counts"])
    np.testing.assert_array_equal(np.concatenate(bbs), g["bb"])
    np.testing.assert_array_equal(np.concatenate(fms), g["fm"])
    np.testing.assert_array_equal(np.concatenate(ams), g["am"])
    np.testing.assert_array_equal(np.concatenate(usbs), g["usb"])


@pytest.mark.parametrize("name", golden_names("bb_s16"))
def test_buffering_independence(name):
    """The stream result must not depend on how the input is cut into buffers."""
    g = load_golden(name)
    a, b = make_bb(g), make_bb(g)
    x = g["x"]
    one = a.process(x)
    cuts = [0, 1, 2, 17, 1000, 1001, 5000, x.shape[0]]
    parts = [b.process(x[s:e]) for s, e in zip(cuts[:-1], cuts[1:])]
    np.testing.assert_array_equal(one, np.concatenate(parts))
    np.testing.assert_array_equal(one, g["bb"])


@pytest.mark.parametrize("name", golden_names("ola_"))
def test_ola_filter_matches_reference(name):
    g = load_golden(name)
    block, Fs = int(g["block"]), float(g["Fs"])
    f = orc.FilterOLA(block, float(g["fmin"]), float(g["fmax"]), Fs)
    np.testing.assert_array_equal(orc.filter_taps(block, min(float(g["fmin"]), float(g["fmax"])),
                                                  max(float(g["fmin"]), float(g["fmax"])), Fs), g["taps"])
    np.testing.assert_array_equal(f.kern, g["kern"])
    x = g["x"].view(np.complex64).reshape(-1) if g["x"].dtype != np.complex64 else g["x"]
    out = f.process(x)
    np.testing.assert_array_equal(out, g["out"])
    # independent time-domain identity
    td = orc.filter_timedomain_f64(g["taps"], x)
    assert rel_rms(out, td) < 5e-6


def test_fast_atan2_matches_formula():
    rng = np.random.default_rng(1)
    for a, b in rng.integers(-32768, 32768, size=(2000, 2)):
        v = orc.fast_atan2_i32(a, b)
        if a == 0 and b == 0:
            assert v == 0
            continue
        aabs = abs(int(a))
        if b >= 0:
            ang = 4096 - int(np.trunc(4096 * (int(b) - aabs) / (int(b) + aabs)))
        else:
            ang = 12288 - int(np.trunc(4096 * (int(b) + aabs) / (aabs - int(b))))
        ang = ang if a >= 0 else -ang
        assert v == np.int16(ang)


def test_shift_operators_floor():
    """test/coretest.cc:10-25 -- >> on negatives is an arithmetic (floor) shift."""
    assert (np.int32(-128) >> 1) == -64 and (np.int32(128) >> 1) == 64
    assert (np.int32(-1) >> 14) == -1


def test_fft_stand_in_against_numpy():
    rng = np.random.default_rng(0)
    for n in (8, 64, 96, 8192):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        assert rel_rms(orc.fft_f64(x, +1), np.fft.fft(x)) < 1e-13
        assert rel_rms(orc.fft_f64(x, -1), np.fft.ifft(x) * n) < 1e-13


@pytest.mark.parametrize("name", ["cast_cu8", "cast_cs8"])
def test_autocast_matches_reference(name):
    g = load_golden(name)
    np.testing.assert_array_equal(orc.autocast_cs16(g["x"]), g["out"])


@pytest.mark.parametrize("name", golden_names("deemph_"))
def test_fmdeemph_matches_reference(name):
    g = load_golden(name)
    d = orc.FMDeemph(float(g["Fs"]))
    bs = int(g["buffer_size"])
    out = np.concatenate([d.process(g["x"][o:o + bs]) for o in range(0, g["x"].shape[0], bs)])
    np.testing.assert_array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names("rbb_"))
def test_real_baseband_matches_reference(name):
    g = load_golden(name)
    o = orc.BaseBand(float(g["Fc"]), float(g["Ff"]), float(g["width"]), int(g["order"]), int(g["sub_sample"]))
    o.config(float(g["Fs"]), int(g["buffer_size"]))
    assert o.lut_inc == int(g["ref_lut_inc"])
    np.testing.assert_array_equal(o.kernel_i32(), g["ref_kernel"])
    bs = int(g["buffer_size"])
    outs = [o.process(g["x"][k:k + bs]) for k in range(0, g["x"].shape[0], bs)]
    np.testing.assert_array_equal(np.array([y.shape[0] for y in outs], dtype=np.uint32), g["counts"])
    np.testing.assert_array_equal(np.concatenate(outs), g["bb"])


def test_autocast_table_oracle_reproduces_reference():
    """oracle.autocast (numpy) == the reference's AutoCast<Scalar> on every pair of its table; pairs the reference
    refuses are refused (tests/golden/cast_table.npz, generated by gen_golden.py from the live reference)."""
    g = load_golden("cast_table")
    n = 0
    for k in g.files:
        if not k.startswith("y_"):
            continue
        _, i, o = k.split("_")
        assert orc.autocast_supported(int(i), int(o))
        np.testing.assert_array_equal(orc.autocast(g["x"], int(i), int(o)), g[k], err_msg=k)
        n += 1
    assert n == 24 and len(g["refused"]) == 24
    for i, o in g["refused"]:
        assert not orc.autocast_supported(int(i), int(o))


@pytest.mark.parametrize("name", golden_names("rbb8_"))
def test_real_baseband_int8_oracle_reproduces_reference(name):
    """BaseBand<int8_t> (16-bit arithmetic throughout, asymmetric complex<int16_t> division): oracle == reference."""
    g = load_golden(name)
    o = orc.BaseBand(float(g["Fc"]), float(g["Ff"]), float(g["width"]), int(g["order"]), int(g["sub_sample"]), scalar=orc.S8)
    bs = int(g["buffer_size"])
    o.config(float(g["Fs"]), bs)
    np.testing.assert_array_equal(o.kernel_i32(), g["ref_kernel"])
    assert o.lut_inc == int(g["ref_lut_inc"])
    x = g["x"]
    outs = [o.process(x[k:k + bs]) for k in range(0, x.shape[0], bs)]
    np.testing.assert_array_equal(np.array([y.shape[0] for y in outs], dtype=np.uint32), g["counts"])
    np.testing.assert_array_equal(np.concatenate(outs), g["bb"])
