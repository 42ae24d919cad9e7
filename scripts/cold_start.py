#!/usr/bin/env python
"""
What `tesseroid.gz(lon, lat, height, model, njobs=N)` costs from Python model objects to the host
result, first and second call of a fresh process, on the BASELINE configs[3] problem (10^6
tesseroids x the 10^6-point grid).  With TVD_DEBUG_TIMING=1 libtvd prints where the set-up time
goes (model creation stages, per-device context / replication stages, prepare and compute phases).

    [TVD_DEBUG_TIMING=1] python scripts/cold_start.py [--gpus N]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from tesseroid_variable_density_b200 import tesseroid  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    lons, lats, heights = bench.c4_points()
    model = bench.c4_model_object()
    pairs = float(model.size) * lons.size
    for label in ("first call (CUDA initialised inside: ~1 s per GPU of driver time)",
                  "second call (model flattened, discretised and replicated again)",
                  "third call"):
        t0 = time.perf_counter()
        tesseroid.gz(lons, lats, heights, model, ratio=2.5, delta=0.1, njobs=a.gpus)
        dt = time.perf_counter() - t0
        print("tesseroid.gz(..., njobs=%d) %-70s %7.3f s  %.3e evals/s" % (a.gpus, label, dt, pairs / dt),
              flush=True)


if __name__ == "__main__":
    main()
