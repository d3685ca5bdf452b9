"""Worker of tests/test_parallel_gloo.py (lives in the package so that spawned processes can import it)."""
import torch
import torch.distributed as dist

from . import parallel


def worker(rank, world, port, channels, n, q):
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        lo, hi = parallel.channel_range(rank, world, channels)
        # what a rank's bank would produce: a deterministic function of (channel, output index)
        ch = torch.arange(lo, hi, dtype=torch.int32).view(-1, 1)
        local = (ch * 1000 + torch.arange(n, dtype=torch.int32).view(1, -1)).to(torch.int16)
        full = parallel.gather_channel_outputs(local, channels)
        expect = (torch.arange(channels, dtype=torch.int32).view(-1, 1) * 1000
                  + torch.arange(n, dtype=torch.int32).view(1, -1)).to(torch.int16)
        ok = bool(torch.equal(full, expect))
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # bench.py's reduction: max over ranks
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok, float(t.item())))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))
