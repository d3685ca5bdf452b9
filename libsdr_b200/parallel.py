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


# ---- peer windows: the gather without a collective (sdrg_peer_*, include/sdrg.h) -----------------------
class PeerWindow:
    """`nbytes` of rank `root`'s HBM that every rank of the group can address (CUDA IPC over NVLink).

    Producers pass addresses inside the window as the output pointers of the *_process_dev calls, so
    the finalize kernels store the demodulated samples straight into the consumer's memory -- no
    collective kernel competes with the compute kernels for SMs.  Layout of the first 4 KiB:
      slots[0..world)  one 64-bit progress flag per producer rank   (offset 0)
      ack              the consumer's progress                       (offset 2048)
    data starts at `data_offset` = 4096.  `ok` is False (on every rank alike) when any rank could not
    map the window; callers then fall back to the NCCL gather.
    """
    HEADER = 4096
    ACK_OFFSET = 2048

    def __init__(self, nbytes, root=0, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        self._lib, self._C = _lib, C
        self.root, self.group = root, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = int(nbytes) + self.HEADER
        self.base = None
        self.owner = self.rank == root
        handle = [None]
        err = 0
        if self.owner:
            p, hd = C.c_void_p(), (C.c_ubyte * 64)()
            try:
                _lib.call("sdrg_peer_window_create", self.nbytes, C.byref(p), hd)
                self.base, handle[0] = p.value, bytes(hd)
            except Exception:
                err = 1
        dist.broadcast_object_list(handle, src=root, group=group)
        if not self.owner and handle[0] is not None:
            p = C.c_void_p()
            try:
                _lib.call("sdrg_peer_window_open", (C.c_ubyte * 64).from_buffer_copy(handle[0]), C.byref(p))
                self.base = p.value
            except Exception:
                err = 1
        elif handle[0] is None:
            err = 1
        t = torch.tensor([err], dtype=torch.int32, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        self.ok = int(t.item()) == 0
        if not self.ok:
            self.close()

    @property
    def data_offset(self):
        return self.HEADER

    def ptr(self, offset=0):
        return self._C.c_void_p(self.base + int(offset))

    def slot(self, rank):
        return self._C.c_void_p(self.base + 8 * rank)

    def ack(self):
        return self._C.c_void_p(self.base + self.ACK_OFFSET)

    def signal(self, slot_ptr, value, stream):
        self._lib.call("sdrg_peer_signal", slot_ptr, int(value), stream)

    def wait(self, slot_ptr, n_slots, value, stream, timeout_ms=10000):
        self._lib.call("sdrg_peer_wait", slot_ptr, int(n_slots), int(value), int(timeout_ms), stream)

    def timed_out(self):
        f = self._C.c_int(0)
        self._lib.call("sdrg_peer_wait_timed_out", self._C.byref(f))
        return bool(f.value)

    def read(self, offset, nbytes):
        """Owner only: copy window bytes to a numpy array (synchronous)."""
        import numpy as np
        out = np.empty(int(nbytes), dtype=np.uint8)
        self._lib.call("sdrg_memcpy_d2h_async", out.ctypes.data_as(self._C.c_void_p), self.ptr(offset), int(nbytes), None)
        self._lib.call("sdrg_device_synchronize")
        return out

    def close(self):
        if self.base is None:
            return
        try:
            self._lib.call("sdrg_peer_window_destroy" if self.owner else "sdrg_peer_window_close", self._C.c_void_p(self.base))
        except Exception:
            pass
        self.base = None


def bank_window_layout(channels, stride, kinds=("fm", "am"), elem_bytes=2, slots=2):
    """Byte offsets (relative to PeerWindow.data_offset) of the gathered (channels, stride) arrays of a
    sharded bank: `slots` generations (double buffering) x kinds.  Returns ({(slot, kind): offset}, total)."""
    off, table = 0, {}
    row = (channels * stride * elem_bytes + 255) // 256 * 256
    for s in range(slots):
        for k in kinds:
            table[(s, k)] = off
            off += row
    return table, off
