#!/usr/bin/env python
"""
Replay ALL stored results of the reference's validation scripts
(tests/golden/reference_results = code/scripts/results of the reference: 450 shell sweeps and
the 6 Neuquen grids) through the CUDA path in one command, the role `make results` plays in the
reference (code/Makefile:70-72).  Prints, per family, how far the recomputed numbers are from
the stored ones.

    python scripts/replay_goldens.py [--engine cuda|oracle] [--family linear ...]
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import cases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="cuda", choices=["cuda", "oracle"])
    ap.add_argument("--family", nargs="*", default=["linear", "exponential", "sine", "grid-search",
                                                    "neuquen"])
    ap.add_argument("--json", default=None, help="write the per-file deviations here")
    ap.add_argument("--write-dir", default=None,
                    help="also write the recomputed results as <dir>/<family>/<file>.npz with the "
                         "reference's own keys (what `make results` produces there)")
    args = ap.parse_args()
    if args.engine == "cuda":
        engine = cases.CudaEngine()
    else:
        import oracle
        engine = cases.OracleEngine(oracle)
    report = {}
    for family in args.family:
        files = sorted(glob.glob(os.path.join(cases.REF_RESULTS, family, "*.npz")))
        t0 = time.time()
        worst = (0.0, None)
        n_exact = n_total = 0
        for path in files:
            fname = os.path.basename(path)
            if family == "neuquen":
                from tesseroid_variable_density_b200 import tesseroid
                kind, field = fname[:-4].split("-")
                stored = np.load(path)
                basin, top, bottom = cases.neuquen_model()
                basin.addprop("density", [cases.neuquen_densities(top, bottom)[kind]] * basin.size)
                if args.engine == "cuda":
                    got = getattr(tesseroid, field)(stored["lon"], stored["lat"], stored["height"],
                                                    basin)
                else:
                    idx = np.linspace(0, stored["lon"].size - 1, 200).astype(int)
                    got = np.full(stored["lon"].size, np.nan)
                    got[idx] = engine(field, stored["lon"][idx], stored["lat"][idx],
                                      stored["height"][idx], basin, cases.DEFAULT_RATIO[field], 1e-2)
                if args.write_dir and args.engine == "cuda":
                    os.makedirs(os.path.join(args.write_dir, family), exist_ok=True)
                    np.savez(os.path.join(args.write_dir, family, fname), result=got,
                             lon=stored["lon"], lat=stored["lat"], height=stored["height"],
                             shape=stored["shape"])
                ok = ~np.isnan(got)
                dev = float(np.max(np.abs(got[ok] - stored["result"][ok])) /
                            np.max(np.abs(stored["result"])))
                n_total += int(ok.sum())
                n_exact += int(np.sum(got[ok] == stored["result"][ok]))
            else:
                got, want = cases.replay_golden(engine, family, fname)
                if args.write_dir:
                    os.makedirs(os.path.join(args.write_dir, family), exist_ok=True)
                    stored = np.load(path)
                    out = {k: stored[k] for k in stored.files if k != "differences"}
                    out["differences"] = got.reshape(stored["differences"].shape)
                    np.savez(os.path.join(args.write_dir, family, fname), **out)
                # the stored numbers are percentages; report their absolute deviation
                dev = float(np.max(np.abs(got - want)))
                n_total += got.size
                n_exact += int(np.sum(got == want))
            report[family + "/" + fname] = dev
            if dev > worst[0]:
                worst = (dev, fname)
        unit = "relative to max|field|" if family == "neuquen" else "percentage points"
        print("%-12s %3d files  %6d values  %5d bit-identical  worst deviation %.3e (%s) in %s  "
              "[%.1f s]" % (family, len(files), n_total, n_exact, worst[0], unit, worst[1],
                            time.time() - t0), flush=True)
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(report, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
