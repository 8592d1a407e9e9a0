"""CPU tier: the N>1 path (pair-index sharding + result gather) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from mdrp_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 10000, 100003):
        for w in (1, 2, 3, 8):
            b = sharding.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and len(b) == w + 1
            sizes = np.diff(b)
            assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


def test_shard_of_ragged_offsets():
    offsets = np.array([0, 5, 5, 12, 20, 21])
    seen = []
    for r in range(2):
        p0, p1, n0, n1, loc = sharding.shard_of(offsets, r, 2)
        assert loc[0] == 0 and loc[-1] == n1 - n0 and len(loc) == p1 - p0 + 1
        seen.append((p0, p1, n0, n1))
    assert seen == [(0, 3, 0, 12), (3, 5, 12, 21)]


def _fake_estimate(offsets, x1, x2, d1, d2, cams):
    """Stand-in for the GPU call: per-pair 'model' = sum of the pair's x1, mask = d1 > 0."""
    from mdrp_b200 import _native as nv
    n = len(offsets) - 1
    models = np.zeros(n, dtype=nv.MODEL_DTYPE)
    stats = np.zeros(n, dtype=nv.STATS_DTYPE)
    for i in range(n):
        models["scale"][i] = x1[offsets[i]:offsets[i + 1]].sum()
        stats["num_inliers"][i] = offsets[i + 1] - offsets[i]
        if cams is not None:
            models["f1"][i] = cams[i, 0]
    return models, stats, (d1 > 0).astype(np.uint8)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 30, size=11)
    offsets = np.r_[0, np.cumsum(sizes)]
    n = int(offsets[-1])
    x1, x2 = rng.normal(size=(n, 2)), rng.normal(size=(n, 2))
    d1, d2 = rng.normal(size=n), rng.normal(size=n)
    cams = np.arange(11 * 8, dtype=np.float64).reshape(11, 8)
    out = sharding.estimate_sharded(_fake_estimate, offsets, x1, x2, d1, d2, cams, rank, world)
    dist.barrier()
    if rank == 0:
        ref = _fake_estimate(offsets, x1, x2, d1, d2, cams)
        ok = all(np.array_equal(a, b) for a, b in zip(out, ref))
        q.put(ok)
    else:
        assert out is None
    dist.destroy_process_group()


def test_two_ranks_gather_in_pair_order():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
