"""Walker sharding + lnL gather on CPU (gloo, world_size 2) — the N>1 host logic of bench.py."""
import os
import socket

import numpy as np
import pytest

from starfish_b200.dist import gather_lnl, shard_range, shard_sizes


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_gather_is_identity_without_process_group():
    import torch

    t = torch.arange(5, dtype=torch.float64)
    assert gather_lnl(t, 5) is t


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_walkers, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_walkers, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float64) * 1.5 - 3.0  # stand-in for this rank's lnL shard
        full = gather_lnl(local, n_walkers)
        q.put((rank, full.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_walkers", [8, 7])
def test_gather_lnl_gloo_world2(n_walkers):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_walkers, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(n_walkers) * 1.5 - 3.0
    for _, full in got:
        assert np.array_equal(full, want)
