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


def make_gpu_compute(prepared, ratio, n_points_total, device=None):
    """
    `compute` callable for :func:`forward_sharded`: evaluates a chunk of points on this rank's
    GPU through ``tvd_forward_device`` with the summation plan of the WHOLE job
    (``tvd_plan_groups(model, n_points_total)``), so that the gathered result is bit-identical
    to a single-GPU run of all points.  `prepared` is a
    :class:`~tesseroid_variable_density_b200.tesseroid.PreparedModel` built on that GPU.
    """
    import ctypes as C

    import torch

    from . import _lib
    from .tesseroid import STACK_SIZE

    lib = _lib.load()
    if device is None:
        device = torch.cuda.current_device()
    n_groups = lib.tvd_plan_groups(prepared.handle, int(n_points_total))

    def compute(field, lon_rad, sinlat, coslat, radius):
        n = int(np.asarray(lon_rad).size)
        dev = torch.device("cuda", device)
        pts = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
               for a in (lon_rad, sinlat, coslat, radius)]
        out = torch.empty(n, dtype=torch.float64, device=dev)
        stats = np.zeros(_lib.NSTATS, dtype=np.int64)
        stream = torch.cuda.current_stream(dev)
        rc = lib.tvd_forward_device(_lib.FIELDS[field], prepared.handle, int(device), n,
                                    pts[0].data_ptr(), pts[1].data_ptr(), pts[2].data_ptr(),
                                    pts[3].data_ptr(), float(ratio), int(STACK_SIZE), n_groups,
                                    out.data_ptr(), _lib.ptr_i64(stats),
                                    C.c_void_p(stream.cuda_stream))
        _lib.check(rc, "tvd_forward_device")
        if stats[3] != 0:       # same RuntimeWarning as the reference (tesseroid.py:227-228)
            import warnings
            from .tesseroid import _WARNING_MSG
            warnings.warn(_WARNING_MSG, RuntimeWarning)
        return out.cpu().numpy()

    return compute
