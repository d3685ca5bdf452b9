"""FilterNode<float> throughput over block sizes (run with PYTHONPATH=.)."""
import torch
from libsdr_b200 import synth
from libsdr_b200.nodes import FilterNode
torch.cuda.set_device(0)
n = 1 << 24
x = torch.from_numpy(synth.c2_input(1 << 20)).cuda().repeat(16, 1).view(torch.complex64).reshape(-1)
for block in (16, 64, 256, 1024, 2048, 4096):
    for nf in (1, 4):
        f = FilterNode(block)
        for k in range(nf):
            f.addFilter(100e3 + 50e3 * k, 300e3 + 50e3 * k)
        f.config(sample_rate=20e6, buffer_size=block)
        for _ in range(2): f.process(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record(); f.process(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[2]
        print("block %5d filters %d: %8.3f ms -> %8.1f MS/s in" % (block, nf, ms, n / ms / 1e3), flush=True)
