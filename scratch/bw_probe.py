import torch, time
torch.cuda.set_device(0)
def timeit(f, n=20, w=5):
    for _ in range(w): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(n):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2], ts[0]
for mb in (512, 2048):
    n = mb*1024*1024//4
    x = torch.randn(n, device='cuda'); y = torch.empty_like(x)
    flush = torch.empty(256*1024*1024//4, device='cuda')
    med,best = timeit(lambda: x.sum())
    print(f"sum {mb}MB: med {med*1e3:.1f} us best {best*1e3:.1f} us -> {mb*1.048576/med:.0f} GB/s (best {mb*1.048576/best:.0f})")
    med,best = timeit(lambda: y.copy_(x))
    print(f"copy {mb}MB: med {med*1e3:.1f} us -> r+w {2*mb*1.048576/med:.0f} GB/s (best {2*mb*1.048576/best:.0f})")
    med,best = timeit(lambda: torch.max(x))
    print(f"max {mb}MB: med {med*1e3:.1f} us -> {mb*1.048576/med:.0f} GB/s")
