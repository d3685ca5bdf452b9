"""Size-independent properties at BASELINE.json's full sizes (where the oracle would take minutes):
chunking invariance, linearity, bank == single node, shift invariance of the FFT filter."""
import numpy as np
import pytest

from conftest import rel_rms
from libsdr_b200 import synth
from libsdr_b200.nodes import IQBaseBand, RxChain, ChannelBank, FilterNode, DEMOD_FM, DEMOD_NONE

pytestmark = pytest.mark.gpu


def _bb(cfg, bs, path=0):
    bb = IQBaseBand(cfg["scalar"], cfg["Fc"], cfg["Ff"], cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"])
    if path:
        bb.setFloatPath(path)
    bb.config(sample_rate=cfg["Fs"], buffer_size=bs)
    return bb


def test_c2_full_size_chunking_and_linearity():
    """Config 2 at 64 buffers of 2^20 complex floats (512 MiB): (i) one launch == 64 launches within
    float round-off, (ii) the folded kernel is linear: f(a x + b y) = a f(x) + b f(y), (iii) folded ==
    direct kernel, all to <= 1e-5 relative RMS; (iv) the count law."""
    import torch
    cfg = dict(synth.C2)
    bs, nb = cfg["buffer_size"], cfg["n_buffers"]
    seg = torch.from_numpy(synth.c2_input(4 << 20)).cuda()
    x = seg.repeat(nb // 4, 1)
    y = torch.roll(x, 12345, dims=0) * 0.5
    one = RxChain(_bb(cfg, bs), DEMOD_FM)
    b1, a1, counts = one.process(x, bs)
    many = RxChain(_bb(cfg, bs), DEMOD_FM)
    parts = [many.process(x[k * bs:(k + 1) * bs], bs) for k in range(nb)]
    b2 = torch.cat([p[0] for p in parts]); a2 = torch.cat([p[1] for p in parts])
    torch.cuda.synchronize()
    assert int(counts.sum()) == (bs * nb - 1) // 416
    assert rel_rms(b2.cpu().numpy(), b1.cpu().numpy()) < 1e-5
    assert rel_rms(a2.cpu().numpy(), a1.cpu().numpy()) < 1e-5
    fx = _bb(cfg, bs).process(x); fy = _bb(cfg, bs).process(y); fxy = _bb(cfg, bs).process(2.0 * x - 3.0 * y)
    torch.cuda.synchronize()
    assert rel_rms(fxy.cpu().numpy(), (2.0 * fx - 3.0 * fy).cpu().numpy()) < 1e-5
    fd = _bb(cfg, bs, path=1).process(x[:8 * bs])
    torch.cuda.synchronize()
    assert rel_rms(fx[:fd.shape[0]].cpu().numpy(), fd.cpu().numpy()) < 1e-5


def test_c4_full_size_bank_equals_single_nodes():
    """Config 4 (256 channels, 100 MS/s int16) over 8 buffers of 2^20: every checked channel of the
    bank is bit-identical to a stand-alone IQBaseBand<int16_t> node on the GPU (which is itself pinned
    against the reference), in one launch and buffer by buffer."""
    import torch
    cfg = dict(synth.C4)
    bs, nb = cfg["buffer_size"], 8
    Fc = synth.bank_frequencies(cfg["channels"], cfg["Fs"])
    seg = torch.from_numpy(synth.bank_input(bs, dict(cfg, channels=16))).cuda()
    x = torch.cat([seg, torch.roll(seg, 777, 0), seg.flip(0), torch.roll(seg, -4321, 0)] * 2)
    bank = ChannelBank("s16", Fc, None, cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"])
    bank.config(sample_rate=cfg["Fs"], buffer_size=bs)
    r = bank.process(x, bs, want=("bb", "fm"))
    torch.cuda.synchronize()
    assert r["bb"].shape[1] == (bs * nb - 1) // 2083
    for c in (0, 77, 128, 255):
        node = IQBaseBand("s16", Fc[c], Fc[c], cfg["width"], cfg["order"], cfg["sub_sample"], cfg["oFs"])
        node.config(sample_rate=cfg["Fs"], buffer_size=bs)
        parts = [node.process(x[k * bs:(k + 1) * bs]).clone() for k in range(nb)]
        torch.cuda.synchronize()
        assert torch.equal(r["bb"][c], torch.cat(parts))


def test_c3_full_size_shift_invariance_and_linearity():
    """Config 3 (block 4096, 1 Mi-sample buffers, 16 buffers): delaying the input by whole blocks
    delays the output identically (time invariance of the overlap-save filter across call
    boundaries) and the filter is linear."""
    import torch
    c = dict(synth.C3)
    n = c["buffer_size"] * 4
    x = torch.from_numpy(synth.c2_input(n)).cuda().view(torch.complex64).reshape(-1)

    def run(sig, pieces):
        f = FilterNode(c["block"]); f.addFilter(c["fmin"], c["fmax"]); f.config(sample_rate=c["Fs"], buffer_size=c["block"])
        outs = [f.process(p)[0] for p in torch.chunk(sig, pieces)]
        torch.cuda.synchronize()
        return torch.cat(outs)

    y1 = run(x, 1); y4 = run(x, 4)
    assert rel_rms(y4.cpu().numpy(), y1.cpu().numpy()) < 1e-6
    d = 3 * c["block"]
    xd = torch.cat([torch.zeros(d, dtype=x.dtype, device=x.device), x[:-d]])
    yd = run(xd, 2)
    assert rel_rms(yd[d:].cpu().numpy(), y1[:-d].cpu().numpy()) < 1e-5
    z = torch.roll(x, 999) * (0.3 + 0.2j)
    assert rel_rms(run(x + z, 1).cpu().numpy(), (y1 + run(z, 1)).cpu().numpy()) < 1e-5


def test_maximum_call_size_crosses_the_2_30_split():
    """One call of more than 2^30 samples (the library splits launches at 2^30, csrc/api.cu) equals the same
    stream cut elsewhere: int8 bit-exact, float (window-pipelined kernel) within tolerance; count law."""
    import torch
    n = (1 << 30) + 70001
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    x8 = torch.randint(-100, 101, (n, 2), dtype=torch.int8, device="cuda", generator=g)
    a = IQBaseBand("s8", 100e3, 100e3, 50e3, 15, 50, 0.0); a.config(sample_rate=2.4e6, buffer_size=1 << 20)
    b = IQBaseBand("s8", 100e3, 100e3, 50e3, 15, 50, 0.0); b.config(sample_rate=2.4e6, buffer_size=1 << 20)
    ya = a.process(x8)
    cut = (1 << 29) + 12345
    yb = torch.cat([b.process(x8[:cut]), b.process(x8[cut:])])
    torch.cuda.synchronize()
    assert ya.shape[0] == (n - 1) // 50 and torch.equal(ya, yb)
    del x8, ya, yb
    torch.cuda.empty_cache()
    seg = torch.from_numpy(synth.c2_input(1 << 22)).cuda()
    xf = seg.repeat(n // seg.shape[0] + 1, 1)[:n].contiguous()
    cfg = dict(synth.C2)
    fa, fb = _bb(cfg, 1 << 20), _bb(cfg, 1 << 20)
    za = fa.process(xf)
    zb = torch.cat([fb.process(xf[:cut]), fb.process(xf[cut:])])
    torch.cuda.synchronize()
    ss = int(cfg["Fs"] / cfg["oFs"])
    assert za.shape[0] == (n - 1) // ss == zb.shape[0]
    e = rel_rms(za.cpu().numpy().astype(np.float64).view(np.complex128), zb.cpu().numpy().astype(np.float64).view(np.complex128))
    assert e < 1e-5, e
