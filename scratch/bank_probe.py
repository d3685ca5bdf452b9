import torch, numpy as np
from libsdr_b200 import synth
from libsdr_b200.nodes import ChannelBank
def timeit(f, n=3, w=2):
    for _ in range(w): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
c=dict(synth.C4); bs=c["buffer_size"]
x=torch.from_numpy(synth.bank_input(bs, dict(c, channels=16))).cuda()
fc=synth.bank_frequencies(c["channels"], c["Fs"])
bank=ChannelBank("s16", fc, None, c["width"], c["order"], c["sub_sample"], c["oFs"]); bank.config(sample_rate=c["Fs"], buffer_size=bs)
n_out=bank.outputs_for(bs)+1
bufs={"fm":torch.zeros((256,n_out),dtype=torch.int16,device="cuda"),"am":torch.zeros((256,n_out),dtype=torch.int16,device="cuda")}
ms=timeit(lambda: bank.process(x, bs, want=("fm","am"), out=bufs))
print(f"C4 256ch: {ms:.3f} ms per {bs} samples -> {bs/ms/1e3:.1f} MS/s in, {256*bs/ms/1e6:.1f} G ch-samples/s")
