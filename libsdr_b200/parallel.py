"""Multi-GPU plumbing of the receive chain: one process per GPU (torch.distributed, NCCL on GPUs,
gloo in the CPU tests).  The hot path has no exchange step: a channel bank is
sharded by contiguous channel ranges, every rank sees the whole input stream, and the only
collective is the gather of the demodulated outputs."""
import numpy as np


def channel_range(rank, world, channels):
    """Contiguous, balanced channel range [lo, hi) owned by `rank` (the first channels % world ranks
    get one more)."""
    base, extra = divmod(channels, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_local_channels(world, channels):
    return -(-channels // world)


def gather_channel_outputs(local, channels, group=None):
    """all_gather of per-channel outputs: `local` is (local_channels, n) on this rank; returns the
    (channels, n) array of the whole bank on every rank.  Ranks with fewer channels are padded to
    the common maximum so that one fixed-size collective suffices."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cmax = max_local_channels(world, channels)
    n = local.shape[1:]
    pad = torch.zeros((cmax,) + tuple(n), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * cmax,) + tuple(n), dtype=local.dtype, device=local.device)
    # neither NCCL nor gloo has a 16-bit integer type: move the bytes
    dist.all_gather_into_tensor(out.view(torch.uint8), pad.view(torch.uint8), group=group)
    parts = []
    for r in range(world):
        lo, hi = channel_range(r, world, channels)
        parts.append(out[r * cmax: r * cmax + (hi - lo)])
    return torch.cat(parts, dim=0)


def bank_frequencies_for_rank(rank, world, all_fc):
    lo, hi = channel_range(rank, world, len(all_fc))
    return np.asarray(all_fc[lo:hi], dtype=np.float64), lo, hi
