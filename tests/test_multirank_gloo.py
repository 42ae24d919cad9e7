"""
The N>1 path on CPU: world_size-2 (and 3) gloo ranks shard the points with the reference's
chunking rule and gather the chunks (tesseroid_variable_density_b200/distributed.py).  The
per-chunk compute is the oracle here (tests may use it; there is no GPU in this environment);
on the GPU box the same test runs with the CUDA path (tests/test_gpu_multi.py).
The property checked is the reference's own: njobs=k is bit-identical to njobs=1.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, dst):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import oracle
    import cases
    from tesseroid_variable_density_b200.distributed import forward_sharded
    from tesseroid_variable_density_b200.model import flatten_model
    from tesseroid_variable_density_b200.tesseroid import _convert_coords
    from tesseroid_variable_density_b200 import gridder
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh, _, _ = cases.exponential_case(1e4, 5)
    flat = flatten_model(mesh)
    lats, lons, heights = gridder.regular((-90, 90, 0, 360), (7, 5), z=2000.)   # 35 points
    cc = _convert_coords(lons, lats, heights)

    def compute(field, lon, sinlat, coslat, radius):
        return oracle.forward(field, flat.bounds, flat.kind, flat.params, 0.1, 2.5, lon, sinlat,
                              coslat, radius)[0]
    res = forward_sharded("gz", *cc, compute=compute, dst=dst)
    if dst is None or rank == dst:
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dst", [(2, None), (2, 0), (3, None)])
def test_sharded_gather_is_bit_identical(tmp_path, oracle_lib, world, dst):
    import cases
    from tesseroid_variable_density_b200.model import flatten_model
    from tesseroid_variable_density_b200.tesseroid import _convert_coords
    from tesseroid_variable_density_b200 import gridder
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), dst), nprocs=world, join=True)
    mesh, _, _ = cases.exponential_case(1e4, 5)
    flat = flatten_model(mesh)
    lats, lons, heights = gridder.regular((-90, 90, 0, 360), (7, 5), z=2000.)
    cc = _convert_coords(lons, lats, heights)
    want = oracle_lib.forward("gz", flat.bounds, flat.kind, flat.params, 0.1, 2.5, *cc)[0]
    ranks = range(world) if dst is None else [dst]
    for r in ranks:
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert np.array_equal(got, want)
