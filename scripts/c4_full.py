#!/usr/bin/env python
"""
BASELINE.json configs[3] at its FULL size in one call: 1e6 tesseroids x the whole 1e6-point grid
(1e12 tesseroid-point pairs), gz, exponential density, on `--gpus` devices, then a parity spot
check of a few points against the oracle (each point costs the CPU ~1.8e6 GLQ evaluations).

    python scripts/c4_full.py [--gpus N] [--z 10000] [--check 16]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import bench  # noqa: E402
from tesseroid_variable_density_b200.tesseroid import PreparedModel, _convert_coords, forward_raw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--z", type=float, default=10e3)
    ap.add_argument("--check", type=int, default=16)
    a = ap.parse_args()
    flat = bench.c4_flat_model()
    lons, lats, heights = bench.c4_points()
    heights = np.full_like(heights, a.z)
    t0 = time.perf_counter()
    pm = PreparedModel(flat, delta=0.1)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    cc = _convert_coords(lons, lats, heights)
    t_conv = time.perf_counter() - t0
    t0 = time.perf_counter()
    res, st = forward_raw("gz", *cc, pm, 2.5, njobs=a.gpus, return_stats=True)
    dt_cold = time.perf_counter() - t0       # includes replicating the model to every GPU
    t0 = time.perf_counter()
    res, st = forward_raw("gz", *cc, pm, 2.5, njobs=a.gpus, return_stats=True)
    dt = time.perf_counter() - t0
    print("first call (replicates the piece table to each GPU, packs records): %.2f s" % dt_cold)
    pairs = float(flat.size) * lons.size
    print("C4 full: %.3e pairs on %d GPU(s): model build %.2f s, coords %.2f s, forward %.2f s "
          "-> %.3e evals/s; near pairs %d, leaves/eval %.4f" % (
              pairs, a.gpus, t_build, t_conv, dt, pairs / dt, st["near_pairs"],
              st["leaves"] / pairs), flush=True)
    if a.check > 0:
        import oracle
        idx = np.random.default_rng(0).choice(lons.size, a.check, replace=False)
        t0 = time.perf_counter()
        o, _ = oracle.forward("gz", flat.bounds, flat.kind, flat.params, 0.1, 2.5,
                              *[c[idx] for c in cc], n_threads=min(a.check, os.cpu_count() or 1))
        print("oracle on %d sampled points: %.1f s; max |gpu - oracle| / max |oracle| = %.3e"
              % (a.check, time.perf_counter() - t0, np.max(np.abs(res[idx] - o)) / np.max(np.abs(o))))
    pm.close()


if __name__ == "__main__":
    main()
