#!/usr/bin/env python
"""
One process per GPU: every rank builds the model on its own GPU, evaluates its chunk of the
points (the reference's `_split_arrays` rule) and the chunks are gathered -- the only
communication.  The gathered field is bit-identical to a single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29500 examples/multi_gpu_torchrun.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tesseroid_variable_density_b200 import ExponentialDensity, TesseroidMesh, gridder  # noqa: E402
from tesseroid_variable_density_b200.constants import G, SI2MGAL  # noqa: E402
from tesseroid_variable_density_b200.distributed import forward_sharded, make_gpu_compute  # noqa: E402
from tesseroid_variable_density_b200.tesseroid import PreparedModel, _convert_coords, forward_raw  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

mesh = TesseroidMesh((-5, 5, -5, 5, 0, -30e3), (1, 100, 100))
mesh.addprop("density", [ExponentialDensity.between(2670, 3300, 0, -30e3, 5)] * mesh.size)
lats, lons, heights = gridder.regular((-6, 6, -6, 6), (300, 300), z=10e3)
cc = _convert_coords(lons, lats, heights)

prepared = PreparedModel(mesh, delta=0.1, device=local)
compute = make_gpu_compute(prepared, 2.5, lons.size, device=local)
gz = forward_sharded("gz", *cc, compute=compute, dst=0)
if rank == 0:
    gz = gz * G * SI2MGAL
    single = forward_raw("gz", *cc, prepared, 2.5, devices=[local]) * G * SI2MGAL
    print("ranks %d, points %d, tesseroids %d: gz %.3f .. %.3f mGal; identical to one GPU: %s"
          % (world, lons.size, mesh.size, gz.min(), gz.max(), np.array_equal(gz, single)))
prepared.close()
dist.destroy_process_group()
