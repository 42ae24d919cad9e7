#!/usr/bin/env python
"""
BASELINE.json configs[4] (SURVEY.md 8d "C5") at its FULL size: a global 0.1 degree computation
grid (1800 x 3600 = 6.48 M points at 260 km) over a 2 M-tesseroid crust (1000 x 2000 cells of
0.18 degree, top 0, seeded random bottom -20 ... -50 km, per-cell sinusoidal density with two
periods over the cell's own thickness, delta = 0.1), potential + gz + gzz in ONE fused far-field
sweep (ratios 1 / 2.5 / 8), points sharded over `--gpus` devices in one process (the reference's
njobs surface, tesseroid.py:179-195).  Then a parity spot check of `--check` points per field
against the oracle (each point costs the CPU 1e7 GLQ evaluations per field).

    python scripts/c5_full.py [--gpus N] [--check 16] [--points 0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from tesseroid_variable_density_b200 import density as D  # noqa: E402
from tesseroid_variable_density_b200.model import FlatModel  # noqa: E402
from tesseroid_variable_density_b200.tesseroid import (PreparedModel, _convert_coords,  # noqa: E402
                                                       forward_raw_fused)

NAMES, RATIOS = ["potential", "gz", "gzz"], [1, 2.5, 8]


def c5_model(nlat=1000, nlon=2000, seed=20190278):
    rng = np.random.default_rng(seed)
    u = rng.random(nlat * nlon)
    dlat, dlon = 180.0 / nlat, 360.0 / nlon
    j = np.tile(np.arange(nlon), nlat)
    i = np.repeat(np.arange(nlat), nlon)
    b = np.empty((nlat * nlon, 6))
    b[:, 0] = dlon * j
    b[:, 1] = b[:, 0] + dlon
    b[:, 2] = -90.0 + dlat * i
    b[:, 3] = b[:, 2] + dlat
    b[:, 4] = 0.0
    b[:, 5] = -(20e3 + 30e3 * u)
    params = np.zeros((nlat * nlon, 5))
    params[:, 0] = 1650.
    params[:, 1] = 2 * np.pi * 2 / (-b[:, 5])
    params[:, 2] = 1650.
    return FlatModel(b, np.full(nlat * nlon, D.SINE, dtype=np.int32), params)


def c5_grid(nlat=1800, nlon=3600, z=260e3):
    lat = -90 + (np.arange(nlat) + 0.5) * 180.0 / nlat
    lon = (np.arange(nlon) + 0.5) * 360.0 / nlon
    LA, LO = np.meshgrid(lat, lon, indexing="ij")
    return LO.ravel(), LA.ravel(), np.full(LO.size, z)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--check", type=int, default=16)
    ap.add_argument("--points", type=int, default=0, help="only the first N grid points (0 = all)")
    a = ap.parse_args()
    t0 = time.perf_counter()
    flat = c5_model()
    lon, lat, h = c5_grid()
    if a.points > 0:
        sel = np.random.default_rng(1).choice(lon.size, a.points, replace=False)
        sel.sort()
        lon, lat, h = lon[sel], lat[sel], h[sel]
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    pm = PreparedModel(flat, delta=0.1)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    cc = _convert_coords(lon, lat, h)
    t_conv = time.perf_counter() - t0
    t0 = time.perf_counter()
    out, st = forward_raw_fused(NAMES, *cc, pm, RATIOS, njobs=a.gpus, return_stats=True)
    dt = time.perf_counter() - t0
    pairs = float(flat.size) * lon.size
    rec = {"config": "C5", "gpus": a.gpus, "tesseroids": flat.size, "pieces": pm.n_pieces,
           "points": int(lon.size), "fields": NAMES, "ratios": RATIOS,
           "host_arrays_s": t_host, "model_build_s": t_build, "convert_coords_s": t_conv,
           "forward_s": dt, "evals_per_s": pairs / dt, "field_evals_per_s": 3 * pairs / dt,
           "piece_pairs_per_s": float(pm.n_pieces) * lon.size / dt,
           "near_pairs": {f: st[f]["near_pairs"] for f in NAMES}}
    print("C5 full: %.3e (tesseroid, point) pairs x 3 fields on %d GPU(s): model build %.2f s "
          "(%d pieces), forward %.2f s (first call: includes replication to the devices) -> "
          "%.3e evals/s (x3 fields = %.3e field evals/s)" % (
              pairs, a.gpus, t_build, pm.n_pieces, dt, pairs / dt, 3 * pairs / dt), flush=True)
    if a.check > 0:
        import oracle
        idx = np.random.default_rng(0).choice(lon.size, a.check, replace=False)
        nt = min(a.check, os.cpu_count() or 1)
        rec["oracle_check"] = {}
        for f, r in zip(NAMES, RATIOS):
            t0 = time.perf_counter()
            o, _ = oracle.forward(f, flat.bounds, flat.kind, flat.params, 0.1, r,
                                  *[c[idx] for c in cc], n_threads=nt)
            rel = float(np.max(np.abs(out[f][idx] - o)) / np.max(np.abs(o)))
            rec["oracle_check"][f] = rel
            print("oracle on %d sampled points, %s: %.1f s; max |gpu - oracle| / max |oracle| = "
                  "%.3e" % (a.check, f, time.perf_counter() - t0, rel), flush=True)
    print(json.dumps(rec))
    pm.close()


if __name__ == "__main__":
    main()
