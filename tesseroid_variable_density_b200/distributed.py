"""
Point sharding over ranks -- one process per GPU.

The reference's only parallelism is data-parallel over computation points: ``njobs`` worker
processes each walk the whole model for a contiguous chunk of the points and the chunks are
concatenated (code/tesseroid_density/tesseroid.py:179-195, ``_split_arrays`` :303-324).  This
module is the same scheme for ``torch.distributed`` ranks: every rank holds the whole piece
table, evaluates its chunk on its own GPU and the chunks are gathered -- there is no collective
on the compute path.  The gather is the only communication (NCCL on GPUs, gloo on CPUs).

Point sharding is exact: a point's value depends on no other point, and the summation plan
(``n_groups``) is derived from the GLOBAL number of points, so sharded results are bit-identical
to single-device results.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_points, world_size, rank):
    """[start, stop) of `rank`'s chunk: ``n = size//nparts``, the last chunk takes the
    remainder (tesseroid.py:318-323)."""
    n = n_points // world_size
    start = rank * n
    stop = n_points if rank == world_size - 1 else start + n
    return start, stop


def forward_sharded(field, lon_rad, sinlat, coslat, radius, compute, group=None, dst=None):
    """
    Evaluate `field` on this rank's chunk with `compute(field, lon, sinlat, coslat, radius)`
    and gather the chunks.  Every rank passes the FULL point arrays (as the reference's
    dispatcher does before it splits them).

    Returns the full result on every rank (``dst=None``, all-gather) or only on rank `dst`.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = int(np.asarray(lon_rad).size)
    start, stop = shard_bounds(n, world, rank)
    local = np.ascontiguousarray(
        compute(field, lon_rad[start:stop], sinlat[start:stop], coslat[start:stop],
                radius[start:stop]), dtype=np.float64)
    # chunks have different lengths (the last one takes the remainder): pad to the longest
    longest = n - (world - 1) * (n // world)
    buf = torch.zeros(longest, dtype=torch.float64)
    buf[:local.size] = torch.from_numpy(local)
    backend = dist.get_backend(group)
    if backend == "nccl":
        buf = buf.cuda()
    if dst is None:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, parts, dst=dst, group=group)
        if rank != dst:
            return None
    out = np.empty(n)
    for r in range(world):
        s, e = shard_bounds(n, world, r)
        out[s:e] = parts[r][:e - s].cpu().numpy()
    return out
