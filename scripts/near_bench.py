#!/usr/bin/env python
"""
Timing of the near-field walk (Phase B, the restatement of rediscretizer,
code/tesseroid_density/_tesseroid.pyx:34-98) on workloads that actually have near-field pairs:

  shell    the 72-tesseroid shell x the 247-point `global` grid at 2 km (BASELINE configs[0]/[1]):
           every pair is near-field, a few pairs are thousands of leaves deep;
  tensor   the same with gzz at its default ratio 8 (the widest subdivision frontier);
  c4_2km   one 75 776-point slab of BASELINE configs[3] at 2 km: millions of shallow pairs.

Prints, per case, the device time of the walk (CUDA events inside libtvd, `tvd_phase_ms`), the
far-field sweep beside it, leaves/s and nodes/s.

    python scripts/near_bench.py [--reps 5]
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import bench  # noqa: E402
import cases  # noqa: E402
from tesseroid_variable_density_b200 import _lib  # noqa: E402
from tesseroid_variable_density_b200.tesseroid import (PreparedModel, _convert_coords,  # noqa: E402
                                                       forward_raw)


def phase(lib, reset):
    out = (C.c_double * 4)()
    lib.tvd_phase_ms(1 if reset else 0, out)
    return [float(v) for v in out]


def run_case(lib, name, pm, cc, field, ratio, reps):
    forward_raw(field, *cc, pm, ratio)                  # warm-up
    phase(lib, True)
    t0 = time.perf_counter()
    for _ in range(reps):
        res, st = forward_raw(field, *cc, pm, ratio, return_stats=True)
    wall = (time.perf_counter() - t0) / reps
    far, near, post, nfar = phase(lib, True)
    far, near, post = far / reps, near / reps, post / reps
    near_leaves = st["leaves"] - (st["nodes"] - st["nonroot_nodes"] - st["near_pairs"])
    near_nodes = st["nonroot_nodes"] + st["near_pairs"]
    print("%-8s %-9s D=%-3g pairs %9d near %9d | walk %8.3f ms  sweep %8.3f ms  after-sweep %8.3f ms"
          "  call %8.3f ms | near nodes %.3e leaves %.3e -> %.3e leaves/s, %.3e nodes/s" % (
              name, field, ratio, st["pieces"] * cc[0].size, st["near_pairs"], near, far, post,
              wall * 1e3, near_nodes, near_leaves, near_leaves / (near * 1e-3) if near > 0 else 0.0,
              near_nodes / (near * 1e-3) if near > 0 else 0.0), flush=True)
    return {"case": name, "field": field, "ratio": ratio, "walk_ms": near, "sweep_ms": far,
            "after_sweep_ms": post, "call_ms": wall * 1e3, "near_pairs": st["near_pairs"],
            "near_nodes": near_nodes, "near_leaves": near_leaves}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-c4", action="store_true")
    a = ap.parse_args()
    lib = _lib.load()
    out = []
    mesh = cases.exponential_case(1e3, 30)[0]
    # model set-up latency (flattening + K1 + tables) for the 72-tesseroid shell
    from tesseroid_variable_density_b200.model import flatten_model
    flat = flatten_model(mesh)
    PreparedModel(flat, delta=0.1).close()
    t0 = time.perf_counter()
    for _ in range(20):
        PreparedModel(flat, delta=0.1).close()
    t_model = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(20):
        PreparedModel(mesh, delta=0.1).close()
    t_model_flat = (time.perf_counter() - t0) / 20
    print("PreparedModel(72 tesseroids, exp density, delta=0.1): %.3f ms from a FlatModel, %.3f ms "
          "from the mesh object" % (t_model * 1e3, t_model_flat * 1e3), flush=True)
    out.append({"case": "model_setup_72", "ms_flat": t_model * 1e3, "ms_mesh": t_model_flat * 1e3})
    lats, lons, heights = cases.GRIDS["global2km"]()
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    for field, ratio in (("gz", 2.5), ("potential", 1), ("gz", 5), ("gzz", 8), ("gy", 2.5)):
        out.append(run_case(lib, "shell", pm, cc, field, ratio, a.reps))
    pm.close()
    lats, lons, heights = cases.GRIDS["global"]()
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    out.append(run_case(lib, "shell_z0", pm, cc, "gz", 2.5, a.reps))
    pm.close()
    if not a.skip_c4:
        flat = bench.c4_flat_model()
        pm = PreparedModel(flat, delta=0.1)
        lons, lats, heights = bench.c4_points()
        P = 75776
        cc = _convert_coords(lons[:P], lats[:P], np.full(P, 2e3))
        out.append(run_case(lib, "c4_2km", pm, cc, "gz", 2.5, 2))
        cc = _convert_coords(lons[:P], lats[:P], np.full(P, 10e3))
        out.append(run_case(lib, "c4_10km", pm, cc, "gz", 2.5, 2))
        pm.close()
    import json
    print(json.dumps(out))


if __name__ == "__main__":
    main()
