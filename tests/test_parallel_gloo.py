"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups check the
channel sharding and the gather layout used by bench.py for the sharded bank (config C5)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from libsdr_b200 import parallel
from libsdr_b200._gloo_selftest import worker


def test_channel_ranges_partition():
    for world in (1, 2, 3, 4, 8):
        for C in (1, 7, 8, 256, 2048, 2049):
            ranges = [parallel.channel_range(r, world, C) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == C
            for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
                assert b == c and b - a >= d - c >= 0
            assert max(b - a for a, b in ranges) == parallel.max_local_channels(world, C)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world,channels", [(2, 8), (2, 7), (3, 8)])
def test_gather_layout_gloo(world, channels):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, channels, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(t == float(world) for _, _, t in res)


def test_bank_window_layout_and_channel_ranges():
    """Host logic of the peer-window gather: the (slot, kind) arrays tile the window without overlap, rows of every
    rank land at disjoint offsets, and the ranges cover all channels exactly once."""
    from libsdr_b200 import parallel
    C_, stride = 2048, 1009
    table, total = parallel.bank_window_layout(C_, stride)
    spans = sorted((off, off + C_ * stride * 2) for off in table.values())
    assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans[:-1], spans[1:])) and spans[-1][1] <= total
    assert all(off % 256 == 0 for off in table.values())
    for world in (1, 2, 3, 4, 8):
        rows = []
        for r in range(world):
            lo, hi = parallel.channel_range(r, world, C_)
            rows.append((lo * stride * 2, hi * stride * 2))
            assert hi - lo in (C_ // world, C_ // world + 1)
        assert rows[0][0] == 0 and rows[-1][1] == C_ * stride * 2
        assert all(a1 == b0 for (_, a1), (b0, _) in zip(rows[:-1], rows[1:]))
    assert parallel.PeerWindow.HEADER >= parallel.PeerWindow.ACK_OFFSET + 8 and parallel.PeerWindow.ACK_OFFSET >= 8 * 64
