"""Multi-rank host logic of the path on CPU (gloo, world_size 2): replicate ranges, rank-layout-independent
seeds, and the final all-gather of per-replicate summaries (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vgsim_b200 import _shard
from vgsim_b200._capi import NSUMMARY


def test_ranges_partition_the_replicates():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            r = [_shard.replicate_range(g, world, total) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        _shard.replicate_range(2, 2, 10)


def test_seeds_depend_on_global_id_only():
    total = 1000
    one = _shard.replicate_seeds(1000, 0, total, batch=3)
    for world in (2, 4, 8):
        parts = [_shard.replicate_seeds(1000, *_shard.replicate_range(g, world, total), batch=3) for g in range(world)]
        assert np.array_equal(np.concatenate(parts), one)
    assert len(set(one.tolist())) == total
    assert not set(one.tolist()) & set(_shard.replicate_seeds(1000, 0, total, batch=4).tolist())
    assert _shard.sweep_point(256 * 5 + 17, 256) == (5, 17)


def _fake_summary(ids):
    # row r = f(global replicate id): what a rank would compute for the replicates it owns
    ids = np.asarray(ids, dtype=np.float64)
    return np.stack([ids * (j + 1) + j for j in range(NSUMMARY)], axis=1)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = _shard.replicate_range(rank, world, total)
        local = torch.from_numpy(_fake_summary(range(lo, hi)))
        out = _shard.gather_summaries(local, world)
        q.put((rank, out.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gather_summaries_world2_gloo(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_summary(range(total))
    for r in range(world):
        assert np.array_equal(got[r], want)
