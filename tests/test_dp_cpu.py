"""Data-parallel host logic on CPU: shard arithmetic and the feature all-gather over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dinov2_b200  # noqa: F401
from dinov2_b200 import dp


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 65, 511):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert dp.shard_range(128, 1, 2) == (64, 128)      # GPU g gets [g*B, (g+1)*B)
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = dp.shard_range(n_items, rank, world)
        # stand-in for the per-rank engine output: feature row i = f(global image index i)
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 5) + 0.25
        full = dp.all_gather_features(local, n_items)
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7])
def test_all_gather_features_gloo(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(n_items, dtype=np.float32)[:, None] * np.ones((1, 5), np.float32) + 0.25
    for r in range(world):
        assert np.array_equal(got[r], want)
