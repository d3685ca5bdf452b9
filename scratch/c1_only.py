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
from libsdr_b200 import _lib as L
from libsdr_b200.nodes import BaseBand
def timeit(f, n=20, w=3):
    for _ in range(w): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device="cpu"); g.manual_seed(0x5D12)
x8 = torch.randint(0, 256, (nb * bs, 2), dtype=torch.uint8, generator=g).cuda()
bb = IQBaseBand("s16", c["Fc"], c["Ff"], c["width"], c["order"], c["sub_sample"], c["oFs"])
bb.setInputType(L.T_CU8); bb.config(sample_rate=c["Fs"], buffer_size=bs)
ch = RxChain(bb, DEMOD_FM)
ms = timeit(lambda: ch.process(x8, bs)); print("cu8 fused: %.1f GS/s" % (nb * bs / ms / 1e6), flush=True)
xr = torch.randint(-32768, 32768, (nb * bs,), dtype=torch.int16, generator=g).cuda()
rb = BaseBand(300e3, 300e3, 50e3, 32, 50); rb.config(sample_rate=c["Fs"], buffer_size=bs)
chr_ = RxChain(rb, DEMOD_FM)
ms = timeit(lambda: chr_.process(xr, bs)); print("real BaseBand<int16> 32 taps: %.1f GS/s" % (nb * bs / ms / 1e6), flush=True)
