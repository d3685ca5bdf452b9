"""FFTPlan<float> batch throughput (run with PYTHONPATH=.)."""
import torch
from libsdr_b200.nodes import FFTPlan
torch.cuda.set_device(0)
for n in (256, 1024, 4096, 8192):
    batch = (1 << 27) // n                       # 1 GiB of complex64
    x = torch.randn((batch * n, 2), device="cuda", dtype=torch.float32).view(torch.complex64).reshape(-1)
    p = FFTPlan(n, FFTPlan.FORWARD)
    y = p(x)
    ref = torch.fft.fft(x.view(batch, n)[:4], dim=1).reshape(-1)
    err = ((y[:4 * n] - ref).abs().pow(2).mean().sqrt() / ref.abs().pow(2).mean().sqrt()).item()
    for _ in range(3): p(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(7):
        e0.record(); p(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[3]
    e0.record(); torch.fft.fft(x.view(batch, n), dim=1); e1.record(); torch.cuda.synchronize()
    e0.record(); torch.fft.fft(x.view(batch, n), dim=1); e1.record(); torch.cuda.synchronize(); cf = e0.elapsed_time(e1)
    print("n=%5d batch=%7d: %.3f ms -> %6.1f GS/s, %6.0f GB/s (r+w)  rel err vs torch.fft %.1e   [cuFFT via torch, for scale: %.3f ms]" %
          (n, batch, ms, batch * n / ms / 1e6, 16 * batch * n / ms / 1e6, err, cf), flush=True)
