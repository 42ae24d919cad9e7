#!/usr/bin/env python
"""
CPU baselines on the host cores of the box ("Timing the reference CPU
path"): a seeded subsample of the C4 workload -- N random tesseroids x N random grid points,
gz, delta = 0.1, ratio = 2.5 -- through

  (i)   the unmodified reference (oracle/_ref), njobs = 1,
  (ii)  the unmodified reference, njobs = os.cpu_count()   (its own multiprocessing path),
  (iii) the reference with a constant density (its pure-C path, no Python callbacks),
  (iv)  the C restatement (oracle/tess_oracle.c), 1 thread and all threads.

Reported baselines only; nothing here is on the product path.

    python scripts/cpu_baselines.py [--n 2000] [--md profiles/cpu_baselines_r1.md]
"""
import argparse
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import bench  # noqa: E402
import build_ref  # noqa: E402
import oracle  # noqa: E402
from tesseroid_variable_density_b200.tesseroid import _convert_coords  # noqa: E402


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--md", default=None)
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(1)
    bounds = bench.c4_model_arrays()
    lons, lats, heights = bench.c4_points()
    ti = rng.choice(bounds.shape[0], a.n, replace=False)
    pi = rng.choice(lons.size, a.n, replace=False)
    b = bounds[ti]
    lon, lat, h = lons[pi].copy(), lats[pi].copy(), heights[pi].copy()
    dens = bench.c4_density()
    ref = build_ref.load()
    from fatiando.mesher import Tesseroid
    model = [Tesseroid(*row, props={"density": dens}) for row in b]
    model_const = [Tesseroid(*row, props={"density": -343.5}) for row in b]
    pairs = a.n * a.n
    rows = []

    def run(label, fn, nthreads):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        rows.append((label, nthreads, dt, pairs / dt))
        print("%-62s %3d threads  %8.2f s  %.3e evals/s" % (label, nthreads, dt, pairs / dt),
              flush=True)
        return out

    # single-core reference on a quarter of the points keeps the run short; rate is per pair
    q = max(1, a.n // 4)
    t0 = time.perf_counter()
    r1 = ref.gz(lon[:q], lat[:q], h[:q], model, delta=0.1)
    dt = time.perf_counter() - t0
    rows.append(("reference (Cython + Python density callable), njobs=1, %d points" % q, 1, dt,
                 a.n * q / dt))
    print("%-62s %3d threads  %8.2f s  %.3e evals/s" % (rows[-1][0], 1, dt, a.n * q / dt), flush=True)
    rn = run("reference, njobs=%d (its multiprocessing path)" % cores,
             lambda: ref.gz(lon, lat, h, model, delta=0.1, njobs=cores), cores)
    assert np.array_equal(rn[:q], r1)           # njobs does not change any value
    run("reference, constant density (pure C path), njobs=1",
        lambda: ref.gz(lon, lat, h, model_const, delta=0.1), 1)
    cc = _convert_coords(lon, lat, h)
    kinds = np.full(a.n, dens.kind, dtype=np.int32)
    params = np.tile(dens.params, (a.n, 1))
    o1 = run("C restatement (oracle), parametric exponential density",
             lambda: oracle.forward("gz", b, kinds, params, 0.1, 2.5, *cc, n_threads=1)[0], 1)
    run("C restatement (oracle), parametric exponential density",
        lambda: oracle.forward("gz", b, kinds, params, 0.1, 2.5, *cc, n_threads=cores)[0], cores)
    from tesseroid_variable_density_b200.constants import G, SI2MGAL
    assert np.max(np.abs(o1 * G * SI2MGAL - rn)) <= 1e-12 * np.max(np.abs(rn))
    if a.md:
        with open(a.md, "w") as fh:
            fh.write("# CPU baselines (scripts/cpu_baselines.py)\n\n")
            fh.write("Host: %s, %d logical cores. Sample: %d random tesseroids x %d random grid "
                     "points of the C4 workload (seed 1), gz, delta=0.1, ratio=2.5.\n\n"
                     % (cpu_model(), cores, a.n, a.n))
            fh.write("| implementation | threads | seconds | evals/s |\n|---|---|---|---|\n")
            for label, nt, dt, rate in rows:
                fh.write("| %s | %d | %.2f | %.3e |\n" % (label, nt, dt, rate))


if __name__ == "__main__":
    main()
