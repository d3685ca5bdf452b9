import torch, numpy as np, ctypes as C, time
from libsdr_b200 import synth, _lib
from libsdr_b200.nodes import FilterNode, IQBaseBand, RxChain, DEMOD_FM
def timeit(f, n=10, w=3):
    for _ in range(w): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
c=synth.C3
nb=16; n=nb*c["buffer_size"]
x=torch.from_numpy(synth.c2_input(1<<20)).cuda().repeat(nb,1).view(torch.complex64).reshape(-1)
for F in (1,4):
    f=FilterNode(c["block"])
    for k in range(F): f.addFilter(c["fmin"]+k*1e5, c["fmax"]+k*1e5)
    f.config(sample_rate=c["Fs"], buffer_size=c["block"])
    ms=timeit(lambda: f.process(x))
    print(f"C3 filters={F}: {ms:.3f} ms per {n} samples -> {n/ms/1e3:.1f} MS/s in, alg bytes {(8+8*F)*n/ms/1e6:.0f} GB/s")
# C1 int16
c=synth.C1
nb=1024; bs=c["buffer_size"]
xi=torch.from_numpy(synth.c1_input(4*bs)).cuda().repeat(nb//4,1)
bb=IQBaseBand("s16",c["Fc"],c["Ff"],c["width"],c["order"],c["sub_sample"],c["oFs"]); bb.config(sample_rate=c["Fs"],buffer_size=bs)
ch=RxChain(bb,DEMOD_FM)
ms=timeit(lambda: ch.process(xi,bs))
print(f"C1 int16: {ms:.3f} ms per {nb*bs} samples -> {nb*bs/ms/1e3:.1f} MS/s, {4.04*nb*bs/ms/1e6:.0f} GB/s")
