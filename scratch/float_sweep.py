"""Float IQBaseBand kernel bandwidth over sub-sampling factors / tap counts (run with PYTHONPATH=.)."""
import sys
import torch
from libsdr_b200 import _lib
from libsdr_b200.nodes import IQBaseBand

torch.cuda.set_device(0)
n = 1 << 27                                  # 1 GiB of cf32
x = torch.randn((n, 2), device="cuda", dtype=torch.float32)
cases = [(int(a.split(",")[0]), int(a.split(",")[1])) for a in sys.argv[1:]] or [(50, 15), (100, 32), (256, 64), (300, 64), (416, 64), (512, 64), (600, 64), (1000, 64), (2083, 64), (4096, 64),
         (20000, 64), (416, 15), (416, 100), (1000, 100)]
for ss, order in cases:
    bb = IQBaseBand("f32", 100e3, 100e3, 12.5e3, order, ss, 0.0)
    import os
    if os.environ.get("FLOAT_PATH"): bb.setFloatPath(int(os.environ["FLOAT_PATH"]))
    bb.config(sample_rate=20e6, buffer_size=n)
    for _ in range(3):
        bb.process(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(7):
        e0.record(); bb.process(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    _lib.profile_enable(True)
    for _ in range(5):
        bb.process(x)
    torch.cuda.synchronize()
    kms, kn = _lib.profile_read(_lib.KERNEL_IQBB_ACCUM)
    fms, fn = _lib.profile_read(_lib.KERNEL_IQBB_FINALIZE)
    _lib.profile_enable(False)
    print("   kernel %d: accumulate %.3f ms = %.1f GB/s, finalize %.3f ms" % (bb.lastFloatKernel(), kms / max(kn, 1), n * 8 / (kms / max(kn, 1)) / 1e6, fms / max(fn, 1)))
    print("ss=%6d L=%3d  %7.3f ms  %7.1f GB/s (whole call incl. finalize)" % (ss, order, ms, n * 8 / ms / 1e6), flush=True)
