"""A/B timing of library variants: python scratch/variant_probe.py build/variants/libsdrg_v0.so ... (one process per variant)."""
import sys, subprocess, os
if len(sys.argv) > 2:
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, lib], env=dict(os.environ, PYTHONPATH="."))
    sys.exit(0)
from libsdr_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from libsdr_b200 import synth
from libsdr_b200.nodes import FilterNode, FFTPlan
def timeit(f, n=40, w=5):
    for _ in range(w): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / n)
    return min(ts)
c = synth.C3
x = torch.from_numpy(synth.c2_input(1 << 20)).cuda().repeat(16, 1).view(torch.complex64).reshape(-1)
n = x.shape[0]
out = [os.path.basename(sys.argv[1])]
for F in (1, 4):
    f = FilterNode(c["block"])
    for k in range(F): f.addFilter(c["fmin"] + k * 4e5, c["fmax"] + k * 4e5)
    f.config(sample_rate=c["Fs"], buffer_size=c["block"])
    ms = timeit(lambda: f.process(x))
    out.append("F=%d %.1f GS/s" % (F, n / ms / 1e6))
xb = torch.randn((1 << 27, 2), device="cuda").view(torch.complex64).reshape(-1)
for nn in (4096, 8192):
    p = FFTPlan(nn, FFTPlan.FORWARD)
    ms = timeit(lambda: p(xb), n=10)
    out.append("fft%d %.3f ms/GiB" % (nn, ms))
print("  ".join(out), flush=True)
