import torch
from libsdr_b200 import synth
from libsdr_b200.nodes import FilterNode, FFTPlan
c = synth.C3
x = torch.from_numpy(synth.c2_input(1 << 20)).cuda().repeat(16, 1).view(torch.complex64).reshape(-1)
f = FilterNode(c["block"]); f.addFilter(c["fmin"], c["fmax"]); f.config(sample_rate=c["Fs"], buffer_size=c["block"])
for _ in range(3):
    f.process(x)
p = FFTPlan(8192, FFTPlan.FORWARD)
xb = torch.randn((1 << 24, 2), device="cuda").view(torch.complex64).reshape(-1)
for _ in range(3):
    p(xb)
torch.cuda.synchronize()
