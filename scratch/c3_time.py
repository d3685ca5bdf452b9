"""C3 timings: FilterNode block 4096 (1 and 4 filters) and FFTPlan 8192 / 4096 (run with PYTHONPATH=.)."""
import torch
from libsdr_b200 import synth
from libsdr_b200.nodes import FilterNode, FFTPlan
c = synth.C3
x = torch.from_numpy(synth.c2_input(1 << 20)).cuda().repeat(16, 1).view(torch.complex64).reshape(-1)


def timed(fn, reps=7, inner=20):
    """median over `reps` of `inner` back-to-back calls between two events (launch gaps excluded, like bench.py)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[len(ts) // 2]


f = FilterNode(c["block"]); f.addFilter(c["fmin"], c["fmax"]); f.config(sample_rate=c["Fs"], buffer_size=c["block"])
ms = timed(lambda: f.process(x))
print("C3 1 filter: %.4f ms  %.1f GS/s" % (ms, x.numel() / ms / 1e6))
f4 = FilterNode(c["block"])
for k in range(4):
    f4.addFilter(c["fmin"] + k * 400e3, c["fmax"] + k * 400e3)
f4.config(sample_rate=c["Fs"], buffer_size=c["block"])
ms = timed(lambda: f4.process(x))
print("C3 4 filters: %.4f ms  %.1f GS/s of input" % (ms, x.numel() / ms / 1e6))
xb = torch.randn((1 << 27, 2), device="cuda").view(torch.complex64).reshape(-1)
for n in (8192, 4096, 2048, 1024):
    p = FFTPlan(n, FFTPlan.FORWARD)
    ms = timed(lambda: p(xb))
    print("FFT %d: %.4f ms per GiB" % (n, ms))
ms = timed(lambda: torch.fft.fft(xb.view(-1, 8192)))
print("cuFFT 8192 (torch.fft): %.4f ms per GiB" % ms)
