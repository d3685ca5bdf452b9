"""GPU parity of the per-window float kernel (libsdr_b200/csrc/iqbb_fold_perwin.cu): short decimation windows,
one thread group per window, against the oracle (<= 1e-5 relative RMS, north_star's float tolerance).
Covers every threads-per-window class (G = 1..16), both window-border regimes (ss == order-1, ss >> order),
positive / negative / zero / not exactly representable shifts, ragged call cuts that move the first whole
window (d_lo = 1 or 2) and calls too short to hold one (which fall back to the per-sample kernels)."""
import numpy as np
import pytest

from conftest import rel_rms
from libsdr_b200 import synth
from libsdr_b200.nodes import IQBaseBand, RxChain, DEMOD_FM
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

FLOAT_TOL = 1e-5
PERWIN = 5     # sdrg_iqbb_last_float_kernel


def _pair(order, ss, Fc, bs):
    Ff = Fc * 0.96
    g = IQBaseBand("f32", Fc, Ff, 200e3, order, ss, 0.0)
    g.config(sample_rate=20e6, buffer_size=bs)
    o = orc.IQBaseBand(orc.F32, Fc, Ff, 200e3, order, ss, 0.0)
    o.config(20e6, bs)
    return g, o


def _table_fits(order, ss):
    """The per-window kernel is used when the V table (<= ss rows, api.cu upload_fold_tables) leaves room for two tile
    buffers for each of four 256-thread groups (half-warp mapping, ss >= 23), resp. fits at all (thread per window)."""
    table = ss * ((ss + order - 1) | 1) * 8
    if ss <= 22:
        return table <= 160 * 1024
    tile = (order - 1 + 16 * ss + 4) * 8
    return table + 8 * tile <= 220 * 1024


@pytest.mark.parametrize("ss,order", [(2, 2), (2, 3), (3, 4), (7, 8), (8, 9), (15, 15), (16, 15), (16, 17), (17, 5), (24, 25),
                                      (31, 15), (32, 15), (33, 34), (48, 20), (50, 15), (63, 33), (64, 32), (64, 65), (65, 64),
                                      (80, 32), (96, 64), (100, 15), (127, 100), (128, 20), (128, 64), (200, 30), (256, 16), (256, 257)])
@pytest.mark.parametrize("Fc", [1.25e6, -1.25e6, 0.0, 333e3], ids=["pos", "neg", "zero", "frac"])
def test_per_window_kernel_geometries(ss, order, Fc):
    n = 300000
    x = synth.iq_f32(n, 20e6, [(0.5, Fc + 2e3, 0.3), (0.3, -4e6, 1.0)], 0.02, 11)
    g, o = _pair(order, ss, Fc, n)
    cuts = [0, 1, 2, 2 + ss // 2, 3 + 2 * ss, 4096, 4097, 4097 + ss, 50000, 50001 + 3 * ss, 200001, n]
    ys, os_, kernels = [], [], []
    for s, e in zip(cuts[:-1], cuts[1:]):
        ys.append(g.process(x[s:e])); os_.append(o.process(x[s:e]))
        kernels.append(g.lastFloatKernel())
        assert ys[-1].shape == os_[-1].shape
    y, ob = np.concatenate(ys), np.concatenate(os_)
    e = rel_rms(y.astype(np.float64).view(np.complex128), ob.astype(np.float64).view(np.complex128))
    assert e < FLOAT_TOL, (e, kernels)
    if _table_fits(order, ss):
        assert kernels[-1] == PERWIN and kernels[-2] == PERWIN, kernels      # the long calls take the per-window kernel


def test_per_window_kernel_one_call_equals_many():
    """Chunking invariance at a C1-like float shape (15 taps, ss = 50): one call of 2^22 samples against 64 ragged ones."""
    import torch
    n = 1 << 22
    x = synth.iq_f32(n, 2.4e6, [(0.5, 103e3, 0.0), (0.2, 500e3, 1.0)], 0.01, 5)
    xd = torch.from_numpy(x).cuda()
    a = IQBaseBand("f32", 100e3, 100e3, 12.5e3, 15, 50, 0.0); a.config(sample_rate=2.4e6, buffer_size=n)
    b = IQBaseBand("f32", 100e3, 100e3, 12.5e3, 15, 50, 0.0); b.config(sample_rate=2.4e6, buffer_size=n)
    ya = a.process(xd)
    assert a.lastFloatKernel() == PERWIN
    rng = np.random.default_rng(3)
    cuts = np.unique(np.concatenate([[0, n], rng.integers(1, n, 63)]))
    yb = torch.cat([b.process(xd[s:e]) for s, e in zip(cuts[:-1], cuts[1:])])
    torch.cuda.synchronize()
    assert ya.shape == yb.shape
    ya, yb = ya.cpu().numpy().astype(np.float64), yb.cpu().numpy().astype(np.float64)
    assert rel_rms(yb.view(np.complex128), ya.view(np.complex128)) < 2e-6
    o = orc.IQBaseBand(orc.F32, 100e3, 100e3, 12.5e3, 15, 50, 0.0); o.config(2.4e6, n)
    ob = o.process(x[:1 << 20]).astype(np.float64)
    assert rel_rms(ya[:ob.shape[0]].view(np.complex128), ob.view(np.complex128)) < FLOAT_TOL


def test_per_window_kernel_fused_chain():
    """Through the RxChain (FM fused into finalize) with per-buffer segments."""
    bs, nb = 65536, 3
    x = synth.iq_f32(bs * nb, 2.4e6, [(0.5, 103e3, 0.0), (0.2, -300e3, 1.0)], 0.01, 7)
    g = IQBaseBand("f32", 100e3, 100e3, 30e3, 21, 24, 0.0); g.config(sample_rate=2.4e6, buffer_size=bs)
    o = orc.IQBaseBand(orc.F32, 100e3, 100e3, 30e3, 21, 24, 0.0); o.config(2.4e6, bs)
    chain = RxChain(g, DEMOD_FM)
    ofm = orc.FMDemod(orc.F32)
    yb, ya, counts = chain.process(x[:3 * bs], bs)
    assert g.lastFloatKernel() == PERWIN
    obs = [o.process(x[k * bs:(k + 1) * bs]) for k in range(3)]
    oas = [ofm.process(b, inplace=True) for b in obs]
    assert rel_rms(yb.astype(np.float64).view(np.complex128), np.concatenate(obs).astype(np.float64).view(np.complex128)) < FLOAT_TOL
    assert rel_rms(ya, np.concatenate(oas)) < FLOAT_TOL


@pytest.mark.parametrize("seed", range(48))
def test_per_window_kernel_random_configs(seed):
    """Seeded random geometries (ss 2..256, 1..ss+1 taps, any shift), random call cuts, device input at odd sample
    offsets (8-byte aligned pointers: the bulk-copy path's unaligned first / last sample)."""
    import torch
    rng = np.random.default_rng(1000 + seed)
    ss = int(rng.integers(2, 257))
    order = int(rng.integers(1, min(ss + 1, 200) + 1))
    Fs = 20e6
    Fc = float(rng.choice([0.0, 1.25e6, -2.5e6, rng.uniform(-9e6, 9e6)]))
    n = 150000 + int(rng.integers(0, 4096))
    x = synth.iq_f32(n + 1, Fs, [(0.5, Fc + 3e3, 0.3), (0.3, -4e6, 1.0)], 0.02, 100 + seed)
    xd = torch.from_numpy(x).cuda()[1:]                 # odd element offset
    x = x[1:]
    g, o = _pair(order, ss, Fc, n)
    cuts = np.unique(np.concatenate([[0, n], rng.integers(1, n, 4)]))
    ys, os_ = [], []
    for s_, e_ in zip(cuts[:-1], cuts[1:]):
        ys.append(g.process(xd[s_:e_]).cpu().numpy()); os_.append(o.process(x[s_:e_]))
        assert ys[-1].shape == os_[-1].shape
    y, ob = np.concatenate(ys), np.concatenate(os_)
    e = rel_rms(y.astype(np.float64).view(np.complex128), ob.astype(np.float64).view(np.complex128))
    assert e < FLOAT_TOL, (e, ss, order, Fc)
