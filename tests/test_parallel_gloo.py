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
