"""Multi-GPU data parallelism over image pairs.

Pairs are independent (no cross-pair state, per-pair RNG seeded from opt.seed), so the batch is
partitioned by pair index across ranks, one process per GPU.  No data-path collective exists;
torch.distributed is only plumbing: a barrier around timed regions and a gather of the small
result structs (96 B model + 40 B stats + N-byte mask per pair) to rank 0.
"""
import numpy as np


def shard_bounds(n_pairs: int, world_size: int):
    """Contiguous, balanced partition: rank r owns pairs [b[r], b[r+1])."""
    base, rem = divmod(n_pairs, world_size)
    b = [0]
    for r in range(world_size):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def shard_of(offsets, rank: int, world_size: int):
    """(pair_start, pair_end, point_start, point_end, local offsets) of this rank's shard."""
    offsets = np.asarray(offsets, dtype=np.int64)
    b = shard_bounds(len(offsets) - 1, world_size)
    p0, p1 = b[rank], b[rank + 1]
    return p0, p1, int(offsets[p0]), int(offsets[p1]), offsets[p0:p1 + 1] - offsets[p0]


def estimate_sharded(estimate_fn, offsets, x1, x2, d1, d2, cams, rank: int, world_size: int, gather: bool = True):
    """Run `estimate_fn(offsets, x1, x2, d1, d2, cams) -> (models, stats, masks)` on this rank's
    shard and gather the results on rank 0 in pair order (other ranks get None)."""
    p0, p1, n0, n1, loc = shard_of(offsets, rank, world_size)
    res = estimate_fn(loc, x1[n0:n1], x2[n0:n1], d1[n0:n1], d2[n0:n1], None if cams is None else cams[p0:p1])
    if world_size == 1 or not gather:
        return res
    import torch.distributed as dist
    bucket = [None] * world_size if rank == 0 else None
    dist.gather_object(res, bucket, dst=0)
    if rank != 0:
        return None
    models = np.concatenate([b[0] for b in bucket])
    stats = np.concatenate([b[1] for b in bucket])
    masks = np.concatenate([b[2] for b in bucket])
    return models, stats, masks
