"""GPU parity of the FFT plan and the FFT-convolution FilterNode (config C3).  Float path:
tolerance <= 1e-5 relative RMS against (1) the golden vectors produced by the reference's own
FilterSink/FilterSource classes, (2) the oracle, (3) the independent time-domain convolution."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_rms
from libsdr_b200 import _lib, synth
from libsdr_b200.nodes import FFTPlan, FilterNode, Config, ConfigError
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_fft_plan_against_numpy(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64)
    fwd, bwd = FFTPlan(n, FFTPlan.FORWARD), FFTPlan(n, FFTPlan.BACKWARD)
    X = fwd(x)
    assert rel_rms(X, np.fft.fft(x.astype(np.complex128), axis=1)) < 2e-6
    y = bwd(X)                                   # unnormalised, like FFTW
    assert rel_rms(y, x.astype(np.complex128) * n) < 2e-6
    assert rel_rms(bwd(x), np.fft.ifft(x.astype(np.complex128), axis=1) * n) < 2e-6
    # the oracle's stand-in agrees too
    assert rel_rms(X[0], orc.fft_f32(x[0], +1)) < 2e-6


def test_fft_plan_errors():
    with pytest.raises(ConfigError):
        FFTPlan(0, FFTPlan.FORWARD)             # empty buffer (fftplan_fftw3.hh:93-97)
    with pytest.raises(ConfigError):
        FFTPlan((1 << 24) + 1, FFTPlan.FORWARD)  # documented upper limit of the device plan


@pytest.mark.parametrize("n", [1, 3, 5, 6, 7, 12, 96, 100, 127, 1000, 1023, 4095, 4097, 5000, 10000, 44100, 100003,
                               16384, 32768, 65536, 1 << 18, 1 << 20])
def test_fft_plan_any_size(n):
    """FFTPlan<float> accepts any size like the reference's FFTW plan (fftplan_fftw3.hh:83-106): powers of two above 8192
    run the four-step decomposition, everything else Bluestein's convolution on the power-of-two engine."""
    rng = np.random.default_rng(n)
    batch = 3 if n <= 65536 else 2
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    fwd, bwd = FFTPlan(n, FFTPlan.FORWARD), FFTPlan(n, FFTPlan.BACKWARD)
    X = fwd(x)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    assert rel_rms(X, ref) < 5e-6, rel_rms(X, ref)
    y = bwd(X)
    assert rel_rms(y, x.astype(np.complex128) * n) < 1e-5
    assert rel_rms(bwd(x), np.fft.ifft(x.astype(np.complex128), axis=1) * n) < 5e-6


@pytest.mark.parametrize("n", [1, 2, 3, 8, 100, 1024, 4097, 65536, 100003])
def test_fft_plan_double(n):
    """FFTPlan<double> (fftplan_fftw3.hh:12-75): double arithmetic on the device."""
    from libsdr_b200.nodes import FFTPlan64
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    X = FFTPlan64(n, FFTPlan64.FORWARD)(x)
    assert rel_rms(X, np.fft.fft(x, axis=1)) < 1e-12
    assert rel_rms(FFTPlan64(n, FFTPlan64.BACKWARD)(X), x * n) < 1e-12
    with pytest.raises(ConfigError):
        FFTPlan64(0, 0)


@pytest.mark.parametrize("block", [1000, 3000, 4095, 5000, 8192, 16384, 20000])
def test_filter_any_block_size(block):
    """FilterNode(block) for blocks that are not a power of two or exceed 4096 (filternode.hh:236 takes any): same taps,
    same normalisation, same causal convolution -- checked against the time-domain identity and, where
    its O(n^2) DFT is affordable, the oracle's restatement of FilterSink/FilterSource."""
    Fs = 2.4e6
    nblk = 5
    x = synth.iq_f32(nblk * block + 777, Fs, [(0.5, 150e3, 0.0), (0.3, -400e3, 1.0), (0.2, 900e3, 2.0)], 0.01, block).view(np.complex64).reshape(-1)
    f = FilterNode(block)
    f.addFilter(100e3, 200e3); f.addFilter(-500e3, -300e3)
    f.config(sample_rate=Fs, buffer_size=block)
    taps = f.design(0)[1] if block <= 8192 else None
    y1 = f.process(x[:2 * block + 100]); y2 = f.process(x[2 * block + 100:])        # ragged: the remainder waits for the next call
    y = np.concatenate([y1, y2], axis=1)
    assert y.shape == (2, nblk * block)
    for k, (lo, hi) in enumerate([(100e3, 200e3), (-500e3, -300e3)]):
        t = orc.filter_taps(block, lo, hi, Fs)
        if k == 0 and taps is not None:
            np.testing.assert_array_equal(taps, t)
        ref = orc.filter_timedomain_f64(t, x[:nblk * block])
        assert rel_rms(y[k], ref) < TOL, (block, k, rel_rms(y[k], ref))
    if block <= 3000:      # the oracle's own OLA (any size through its O(n^2) DFT) agrees as well
        o = orc.FilterOLA(block, 100e3, 200e3, Fs)
        assert rel_rms(y[0], o.process(x[:nblk * block])) < TOL


def test_fft_device_pointers():
    import torch
    n = 4096
    x = torch.randn(5, n, dtype=torch.complex64, device="cuda")
    X = FFTPlan(n, FFTPlan.FORWARD)(x)
    torch.cuda.synchronize()
    assert rel_rms(X.cpu().numpy(), torch.fft.fft(x.to(torch.complex128)).cpu().numpy()) < 2e-6


@pytest.mark.parametrize("name", golden_names("ola_"))
def test_filter_golden(name):
    g = load_golden(name)
    block, Fs = int(g["block"]), float(g["Fs"])
    f = FilterNode(block)
    idx = f.addFilter(float(g["fmin"]), float(g["fmax"]))
    out_cfg = f.config(sample_rate=Fs, buffer_size=block)
    assert out_cfg.type == _lib.T_CF32 and out_cfg.buffer_size == block
    kern, taps = f.design(idx)
    np.testing.assert_array_equal(taps, g["taps"])          # same float formula as sinc_flt_kernel<float>
    assert rel_rms(kern, g["kern"]) < 1e-6
    x = g["x"].view(np.complex64).reshape(-1)
    y = f.process(x)[0]
    assert y.shape == g["out"].shape
    assert rel_rms(y, g["out"]) < TOL


def test_filter_c3_shape_against_oracle_and_time_domain():
    """Config 3: block 4096 (FFT 8192), band-pass 100..300 kHz at 20 MS/s, the C2 input signal."""
    c = synth.C3
    n = 96 * c["block"]
    x = synth.c2_input(n).view(np.complex64).reshape(-1)
    f = FilterNode(c["block"]); f.addFilter(c["fmin"], c["fmax"]); f.config(sample_rate=c["Fs"], buffer_size=c["block"])
    o = orc.FilterOLA(c["block"], c["fmin"], c["fmax"], c["Fs"])
    y1 = f.process(x[:n // 2])[0]; y2 = f.process(x[n // 2:])[0]     # state carried across calls
    y = np.concatenate([y1, y2])
    ref = o.process(x)
    e = rel_rms(y, ref)
    assert e < TOL, e
    td = orc.filter_timedomain_f64(orc.filter_taps(c["block"], c["fmin"], c["fmax"], c["Fs"]), x[:8 * c["block"]])
    assert rel_rms(y[:8 * c["block"]], td) < TOL


@pytest.mark.parametrize("block", [1024, 4096])       # 4096: the in-place radix-16 bank kernel (fft8k_kernels.cu)
def test_filter_bank_shared_forward_fft(block):
    Fs = 2.4e6
    x = synth.iq_f32(40 * block, Fs, [(0.5, 150e3, 0.0), (0.3, -400e3, 1.0), (0.2, 900e3, 2.0)], 0.01, 5).view(np.complex64).reshape(-1)
    bands = [(100e3, 200e3), (-500e3, -300e3), (1e6, 800e3), (-1.2e6, 1.2e6)]
    f = FilterNode(block)
    for lo, hi in bands:
        f.addFilter(lo, hi)
    f.config(sample_rate=Fs, buffer_size=block)
    y = f.process(x)
    assert y.shape == (4, 40 * block)
    for k, (lo, hi) in enumerate(bands):
        ref = orc.FilterOLA(block, lo, hi, Fs).process(x)
        assert rel_rms(y[k], ref) < TOL
    # retune one filter (FilterSource::setFreq); the stream state carries on
    f.setFreq(0, -200e3, -100e3)
    y2 = f.process(x)
    o = orc.FilterOLA(block, -200e3, -100e3, Fs)
    o.last = None
    ref2 = orc.filter_timedomain_f64(orc.filter_taps(block, -200e3, -100e3, Fs), np.concatenate([x, x]))[40 * block:]
    assert rel_rms(y2[0], ref2) < TOL


def test_filter_rechunks_ragged_input():
    """BufferNode semantics: input sizes need not be multiples of the block size."""
    import torch
    block, Fs = 256, 1e6
    x = synth.iq_f32(10000, Fs, [(0.5, 100e3, 0.0), (0.2, -300e3, 1.0)], 0.01, 9).view(np.complex64).reshape(-1)
    f = FilterNode(block); f.addFilter(50e3, 150e3); f.config(sample_rate=Fs, buffer_size=1000)
    cuts = [0, 1, 255, 256, 257, 1000, 1511, 4096, 4097, 9999, 10000]
    xd = torch.from_numpy(x).cuda()
    parts = []
    for s, e in zip(cuts[:-1], cuts[1:]):
        y = f.process(xd[s:e])
        torch.cuda.synchronize()
        parts.append(y[0].cpu().numpy())
    y = np.concatenate(parts)
    nfull = (10000 // block) * block
    assert y.shape[0] == nfull
    ref = orc.FilterOLA(block, 50e3, 150e3, Fs).process(x[:nfull])
    assert rel_rms(y, ref) < TOL


def test_filter_config_errors():
    f = FilterNode(1024)
    with pytest.raises(ConfigError):
        f.config(Config(_lib.T_CS16, 1e6, 1024, 1))
    assert f.config(Config(_lib.T_CF32, 0.0, 1024, 1)).type == _lib.T_UNDEFINED
    with pytest.raises(ConfigError):
        FilterNode((1 << 22) + 1)
    with pytest.raises(RuntimeError):
        FilterNode(64).process(np.zeros(64, dtype=np.complex64))
