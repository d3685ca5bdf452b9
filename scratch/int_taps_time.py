"""int16 IQBaseBand throughput over tap counts (run with PYTHONPATH=.)."""
import sys
import torch
from libsdr_b200 import synth
from libsdr_b200.nodes import IQBaseBand, RxChain, DEMOD_FM
bs, nb = 1 << 20, 64
xi = torch.from_numpy(synth.c1_input(4 * 65536)).cuda().repeat(nb * bs // (4 * 65536), 1)
cases = [(int(a.split(",")[0]), int(a.split(",")[1])) for a in sys.argv[1:]] or [(15, 50), (32, 50), (33, 50), (48, 416), (64, 416), (64, 50)]
for order, ss in cases:
    bb = IQBaseBand("s16", 100e3, 100e3, 12.5e3, order, ss, 0.0); bb.config(sample_rate=20e6, buffer_size=bs)
    ch = RxChain(bb, DEMOD_FM)
    for _ in range(3):
        ch.process(xi, bs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ch.process(xi, bs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gs = nb * bs / ms / 1e6
    print("int16 %3d taps ss=%4d: %.3f ms -> %6.1f GS/s = %.1f G tap-MAC/s per GPU" % (order, ss, ms, gs, gs * (order - 1)), flush=True)
