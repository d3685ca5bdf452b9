import torch
from libsdr_b200 import synth
from libsdr_b200.nodes import IQBaseBand, RxChain, DEMOD_FM
c = synth.C1
nb, bs = 1024, c["buffer_size"]
xi = torch.from_numpy(synth.c1_input(4 * bs)).cuda().repeat(nb // 4, 1)
for Ff in (c["Ff"], 0.0):
    bb = IQBaseBand("s16", c["Fc"], Ff, c["width"], c["order"], c["sub_sample"], c["oFs"]); bb.config(sample_rate=c["Fs"], buffer_size=bs)
    ch = RxChain(bb, DEMOD_FM)
    for _ in range(3):
        ch.process(xi, bs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ch.process(xi, bs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("C1 Ff=%g: %.3f ms -> %.1f GS/s" % (Ff, ms, nb * bs / ms / 1e6), flush=True)
