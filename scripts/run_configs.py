#!/usr/bin/env python
"""
Run the five configurations of BASELINE.json on the GPU and print one table: accuracy against
the analytical shells / the reference's stored results, time, throughput.

  C1  shell, linear density, gz + potential at 2 km on small grids vs the analytical solution
  C2  shell, exponential and sinusoidal density, all 10 fields, sweeps of D = 1..5, delta = 0.1..1
  C3  Neuquen basin: sediment thickness -> exponential-density model, gz at 10 km (full grid)
  C4  synthetic regional model, 1e6 tesseroids x a 75 776-point slab of the 1e6-point grid, gz
  C5  synthetic global model, 2e6 sine-density tesseroids x a 75 776-point slab of the 0.1 deg
      grid at 260 km, potential + gz + gzz in one fused sweep

C4/C5 are timed on one slab per GPU (the full grids are 13.2 / 85.5 slabs); `--cpu` adds the
unmodified reference (oracle/_ref) on the host cores for C1-C3 (C3 on a sample of its grid).

    python scripts/run_configs.py [--cpu] [--md profiles/configs_r1.md]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import bench  # noqa: E402
import cases  # noqa: E402
from tesseroid_variable_density_b200 import density as D  # noqa: E402
from tesseroid_variable_density_b200 import tesseroid  # noqa: E402
from tesseroid_variable_density_b200.model import FlatModel  # noqa: E402
from tesseroid_variable_density_b200.tesseroid import (PreparedModel, _convert_coords,  # noqa: E402
                                                       forward_raw, forward_raw_fused)

ROWS = []


def row(config, what, pairs, seconds, accuracy, cpu=None):
    ROWS.append((config, what, pairs, seconds, accuracy, cpu))
    rate = pairs / seconds if seconds > 0 else float("nan")
    print("%-3s %-58s pairs %.3e  %.4f s  %.3e evals/s  %s%s" % (
        config, what, pairs, seconds, rate, accuracy,
        "" if cpu is None else "  | reference CPU: %.2f s (%.3e evals/s)" % (cpu, pairs / cpu)),
        flush=True)


def timed(fn, *a, **k):
    fn(*a, **k)                       # warm-up (allocations, caches)
    t0 = time.perf_counter()
    out = fn(*a, **k)
    return out, time.perf_counter() - t0


def c1(ref):
    for T in (1e3, 1e5):
        mesh, _, exact = cases.linear_case(T)
        for gname in ("global2km", "equator2km"):
            lats, lons, heights = cases.GRIDS[gname]()
            ex = exact(heights[0])
            worst, total, cpu = 0.0, 0.0, (0.0 if ref else None)
            for field in ("potential", "gz"):
                res, dt = timed(getattr(tesseroid, field), lons, lats, heights, mesh, delta=None)
                worst = max(worst, cases.percent_difference(res, ex[field]))
                total += dt
                if ref:
                    t0 = time.perf_counter()
                    r = getattr(ref, field)(lons, lats, heights, cases.linear_case(T)[0],
                                            delta=None, njobs=os.cpu_count())
                    cpu += time.perf_counter() - t0
                    assert np.max(np.abs(r - res)) <= 1e-9 * np.max(np.abs(r))
            row("C1", "linear shell T=%g m, %s, potential+gz" % (T, gname), 2 * 72 * lons.size,
                total, "max %.2e %% of analytical" % worst, cpu)


def c2():
    lats, lons, heights = cases.GRIDS["global2km"]()
    for name, mk in (("exponential b=10", lambda: cases.exponential_case(1e4, 10)),
                     ("sine b=2", lambda: cases.sine_case(1e4, 2))):
        mesh, _, exact = mk()
        ex = exact(heights[0])
        worst = {}
        t0 = time.perf_counter()
        n = 0
        for ratio in (1, 2, 3, 4, 5):
            for delta in (0.1, 0.5, 1.0):
                pm = PreparedModel(mesh, delta=delta)
                for field in cases.ALL_FIELDS:
                    r = max(ratio, 1) * (1 if field == "potential" else 1)
                    res = getattr(tesseroid, field)(lons, lats, heights, pm, ratio=r)
                    n += 72 * lons.size
                    if ex[field] != 0 and ratio >= cases.DEFAULT_RATIO[field] and delta <= 0.1:
                        worst[field] = max(worst.get(field, 0.0),
                                           cases.percent_difference(res, ex[field]))
                pm.close()
        dt = time.perf_counter() - t0
        acc = ", ".join("%s %.1e%%" % kv for kv in sorted(worst.items()))
        row("C2", "%s shell, 10 fields x D=1..5 x delta=0.1,0.5,1 (150 runs)" % name, n, dt,
            "at D>=default, delta=0.1: " + acc)


def c3(ref):
    stored = np.load(os.path.join(cases.REF_RESULTS, "neuquen", "exponential-gz.npz"))
    basin, top, bottom = cases.neuquen_model()
    dens = cases.neuquen_densities(top, bottom)["exponential"]
    basin.addprop("density", [dens] * basin.size)
    res, dt = timed(tesseroid.gz, stored["lon"], stored["lat"], stored["height"], basin)
    dev = np.max(np.abs(res - stored["result"])) / np.max(np.abs(stored["result"]))
    cpu = None
    if ref:
        idx = np.linspace(0, stored["lon"].size - 1, 64).astype(int)
        t0 = time.perf_counter()
        r = ref.gz(stored["lon"][idx], stored["lat"][idx], stored["height"][idx], basin,
                   njobs=os.cpu_count())
        cpu = (time.perf_counter() - t0) * stored["lon"].size / idx.size     # extrapolated
        assert np.max(np.abs(r - res[idx])) <= 1e-9 * np.max(np.abs(r))
    row("C3", "Neuquen exponential gz @10 km, 10 647 tesseroids x 6 399 points (incl. model build)",
        basin.size * stored["lon"].size, dt, "max %.1e of the stored result" % dev, cpu)


def c4():
    flat = bench.c4_flat_model()
    pm = PreparedModel(flat, delta=0.1)
    lons, lats, heights = bench.c4_points()
    P = 75776
    for z, label in ((10e3, "10 km"), (2e3, "2 km")):
        cc = _convert_coords(lons[:P], lats[:P], np.full(P, z))
        (res, st), dt = timed(forward_raw, "gz", *cc, pm, 2.5, return_stats=True)
        row("C4", "regional 1e6 tesseroids (%d pieces) x %d-point slab, gz @%s" % (pm.n_pieces, P, label),
            flat.size * P, dt, "near pairs %d, leaves/eval %.4f" % (st["near_pairs"],
                                                                   st["leaves"] / (flat.size * P)))
    pm.close()


def c5():
    nlat, nlon = 1000, 2000
    rng = np.random.default_rng(20190278)
    u = rng.random(nlat * nlon)
    j = np.tile(np.arange(nlon), nlat)
    i = np.repeat(np.arange(nlat), nlon)
    b = np.empty((nlat * nlon, 6))
    b[:, 0] = 0.18 * j
    b[:, 1] = b[:, 0] + 0.18
    b[:, 2] = -90.0 + 0.18 * i
    b[:, 3] = b[:, 2] + 0.18
    b[:, 4] = 0.0
    b[:, 5] = -(20e3 + 30e3 * u)
    params = np.zeros((nlat * nlon, 5))
    params[:, 0] = 1650.
    params[:, 1] = 2 * np.pi * 2 / (-b[:, 5])
    params[:, 2] = 1650.
    flat = FlatModel(b, np.full(nlat * nlon, D.SINE, dtype=np.int32), params)
    t0 = time.perf_counter()
    pm = PreparedModel(flat, delta=0.1)
    build = time.perf_counter() - t0
    lat = -89.95 + 0.1 * np.arange(1800)
    lon = 0.05 + 0.1 * np.arange(3600)
    LA, LO = np.meshgrid(lat, lon, indexing="ij")
    P = 75776
    sel = slice(3_000_000, 3_000_000 + P)
    cc = _convert_coords(LO.ravel()[sel], LA.ravel()[sel], np.full(P, 260e3))
    names, ratios = ["potential", "gz", "gzz"], [1, 2.5, 8]
    (out, st), dt = timed(forward_raw_fused, names, *cc, pm, ratios, return_stats=True)
    row("C5", "global 2e6 sine tesseroids (%d pieces, build %.2f s) x %d-point slab, V+gz+gzz fused"
        % (pm.n_pieces, build, P), flat.size * P, dt,
        "near pairs %s" % [st[f]["near_pairs"] for f in names])
    pm.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--md", default=None)
    args = ap.parse_args()
    ref = None
    if args.cpu:
        import build_ref
        ref = build_ref.load()
    c1(ref)
    c2()
    c3(ref)
    c4()
    c5()
    if args.md:
        with open(args.md, "w") as fh:
            fh.write("# BASELINE.json configurations on one B200 (scripts/run_configs.py)\n\n")
            fh.write("| config | what | pairs | GPU s | GPU evals/s | accuracy | reference CPU s "
                     "(%d cores) |\n|---|---|---|---|---|---|---|\n" % (os.cpu_count() or 1))
            for c, w, n, s, a, cpu in ROWS:
                fh.write("| %s | %s | %.3e | %.4f | %.3e | %s | %s |\n" % (
                    c, w, n, s, n / s, a, "" if cpu is None else "%.2f" % cpu))


if __name__ == "__main__":
    main()
