#!/usr/bin/env python
"""
bench.py -- headline benchmark of the tesseroid forward-modelling hot path.

Metric (BASELINE.json): tesseroid-point evaluations per second, gz, exponential density.
One eval = one (input tesseroid, computation point) pair fully integrated (all its radial
pieces, all its subdivision nodes).

Workload (BASELINE.json configs[3], SURVEY.md 8d "C4"): synthetic regional model of
1000 x 1000 tesseroids of 0.02 deg (lon 280..300, lat -45..-25, top 845 m, seeded random
thickness 200..6000 m), one exponential density (b = 3, -412 kg/m3 at the top to -275 at
-5155 m), delta = 0.1, ratio 2.5, computation points from the 1000 x 1000 grid at 10 km.
A *step* is one pass of the whole model over one batch of `--points` grid points per GPU
(a different slab of the grid every step); ranks own disjoint slabs, so per-GPU work is fixed
("weak").  Inputs are larger than L2 per step only in aggregate; the piece table (~0.4 GB)
is larger than the 126 MB L2 and is streamed once per thread block.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl reference]

Prints ONE JSON line (rank 0).  See the task contract in DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = "tesseroid-point evals/sec (gz, exp density)"
UNIT = "evals/s"
NLAT = NLON = 1000
SEED = 20190277
# FP64 flops of one far-field gz (piece, point) pair, FMA = 2.
#  * F_PAIR_EXECUTED: what the far-field sweep of THIS build executes per pair, from the ncu
#    capture committed with it (profiles/far_sweep_r2.md: 11.4 DADD + 36.7 DMUL + 40.3 DFMA =
#    88.4 instructions = 128.8 flops; round 1: 96.4 instructions = 136.7 flops;
#    smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on x elapsed cycles / pairs).  `roofline.achieved` uses this figure: it is the hardware flop rate of the kernel
#    and can be checked against ncu directly.  It must be re-measured when the kernel changes.
#  * SURVEY.md section 8d (convention v1) ESTIMATED 384 = F_test 76 + F_glq 308 from guessed
#    special-function weights and asked for a one-time ncu calibration on the first kernel; that
#    kernel (profiles/far_sweep_r1a.md) executed 218 flops per pair.
#    Later kernels do the same pairs with fewer operations (cheaper test and cos, angular work
#    shared between pieces), so rates computed with 218 or 384 exceed what the hardware executes;
#    they are reported beside `achieved`, not as `achieved`.
F_PAIR_EXECUTED = 128.8
F_PAIR_CALIBRATED = 218
F_PAIR_V1_ESTIMATE = 384
# static facts of the committed far-field kernel from its ncu capture (profiles/far_sweep_r2.md)
FAR_TRAFFIC_BYTES = 562.7e6
FAR_NCU = {"fp64_pipe_pct_of_peak": 83.9, "issue_active_pct": 59.7,
           "executed_fp64_instructions_per_piece_pair": 88.4,
           "executed_flops_per_piece_pair": 128.8,
           # the same ncu metric for the bare DFMA loop that measures `peak` (profiles/pipe_probe_r2.csv)
           "fp64_pipe_pct_of_peak_bare_dfma_loop": 91.87}


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def c4_model_arrays(n_lat=NLAT, n_lon=NLON):
    """bounds (n, 6) of the C4 synthetic regional model (row-major, latitude slow)."""
    rng = np.random.default_rng(SEED)
    thickness = rng.uniform(200, 6000, n_lat * n_lon)
    j = np.tile(np.arange(n_lon), n_lat)
    i = np.repeat(np.arange(n_lat), n_lon)
    bounds = np.empty((n_lat * n_lon, 6))
    bounds[:, 0] = 280.0 + 0.02 * j
    bounds[:, 1] = bounds[:, 0] + 0.02
    bounds[:, 2] = -45.0 + 0.02 * i
    bounds[:, 3] = bounds[:, 2] + 0.02
    bounds[:, 4] = 845.0
    bounds[:, 5] = 845.0 - thickness
    return bounds


def c4_density():
    from tesseroid_variable_density_b200.density import ExponentialDensity
    return ExponentialDensity.between(-412.0, -275.0, 845.0, -5155.0, 3)


def c4_points():
    from tesseroid_variable_density_b200 import gridder
    lats, lons, heights = gridder.regular((-44.99, -25.01, 280.01, 299.99), (1000, 1000), z=10e3)
    return lons, lats, heights


def c4_flat_model():
    from tesseroid_variable_density_b200.model import FlatModel
    bounds = c4_model_arrays()
    kind, params = c4_density().kind, c4_density().params
    n = bounds.shape[0]
    return FlatModel(bounds, np.full(n, kind, dtype=np.int32), np.tile(params, (n, 1)))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(np.max(smax))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
# reference / CPU baseline arm
# ---------------------------------------------------------------------------------------------
def _reference_callable():
    """(kind, fn): the unmodified reference from oracle/_ref if it was built, else the C port."""
    sys.path.insert(0, os.path.join(HERE, "oracle"))
    import build_ref
    if build_ref.available():
        ref = build_ref.load()
        return "reference", ref
    return "port", None


def cpu_sample_run(n_tess, n_points, n_threads, seed=1):
    """
    Time the reference's CPU implementation on a seeded subsample of the C4 workload
    (n_tess random tesseroids x n_points random grid points), all host threads it can use.
    Returns (evals_per_s, seconds, kind).
    """
    from tesseroid_variable_density_b200.tesseroid import _convert_coords
    rng = np.random.default_rng(seed)
    bounds = c4_model_arrays()
    lons, lats, heights = c4_points()
    ti = rng.choice(bounds.shape[0], n_tess, replace=False)
    pi = rng.choice(lons.size, n_points, replace=False)
    b = bounds[ti]
    dens = c4_density()
    kind, ref = _reference_callable()
    lon, lat, h = lons[pi].copy(), lats[pi].copy(), heights[pi].copy()
    if kind == "reference":
        from fatiando.mesher import Tesseroid            # oracle/shims
        model = [Tesseroid(*row, props={"density": dens}) for row in b]
        t0 = time.perf_counter()
        ref.gz(lon, lat, h, model, njobs=n_threads, delta=0.1)
        dt = time.perf_counter() - t0
    else:
        import oracle
        cc = _convert_coords(lon, lat, h)
        kinds = np.full(n_tess, dens.kind, dtype=np.int32)
        params = np.tile(dens.params, (n_tess, 1))
        t0 = time.perf_counter()
        oracle.forward("gz", b, kinds, params, 0.1, 2.5, *cc, n_threads=n_threads)
        dt = time.perf_counter() - t0
    return n_tess * n_points / dt, dt, kind


def _size_sample(rate, seconds, cores):
    pairs = max(3600.0, rate * seconds)
    n_t = max(60, min(int(np.sqrt(pairs)), 4000))
    n_p = max(cores, min(int(pairs / n_t), 100000))
    return n_t, n_p


def cpu_baseline(target_s=12.0):
    """Bounded CPU sample (about 10-30 s of CPU work in total), all host cores."""
    cores = os.cpu_count() or 1
    # calibrate on a tiny sample (dominated by the pool start-up), then on a 2 s one
    rate, dt, kind = cpu_sample_run(60, 60 * max(1, cores // 8), cores)
    n_t, n_p = _size_sample(rate, 2.0, cores)
    rate, dt, kind = cpu_sample_run(n_t, n_p, cores)
    n_t, n_p = _size_sample(rate, target_s, cores)
    rate, dt, kind = cpu_sample_run(n_t, n_p, cores)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d random tesseroids x %d random grid points of the C4 workload, "
                      "gz, delta=0.1, ratio=2.5, njobs=%d, %.1f s" % (n_t, n_p, cores, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rate, dt, kind = cpu_sample_run(60, 60 * max(1, cores // 8), cores)      # calibration
    n_t, n_p = _size_sample(rate, 2.0, cores)
    rate, dt, kind = cpu_sample_run(n_t, n_p, cores)
    n_t, n_p = _size_sample(rate, 8.0, cores)                                # ~8 s per step
    for w in range(args.warmup):
        cpu_sample_run(n_t, max(cores, n_p // 8), cores, seed=100 + w)
    t_total, evals = 0.0, 0
    for s in range(args.steps):
        r, dt, kind = cpu_sample_run(n_t, n_p, cores, seed=200 + s)
        t_total += dt
        evals += n_t * n_p
    value = evals / t_total
    sample = ("%d random tesseroids x %d random grid points of the C4 workload per step, gz, "
              "delta=0.1, ratio=2.5, njobs=%d" % (n_t, n_p, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "C4 synthetic regional model: 1000x1000 tesseroids of 0.02 deg x "
                        "%d-point slabs of the 1000x1000 grid at 10 km per GPU per step, "
                        "exponential density b=3, gz, ratio=2.5, delta=0.1" % args.points,
            "n_tesseroids": NLAT * NLON, "points_per_gpu_per_step": args.points,
            "l2": "piece table (0.38 GB) exceeds L2; a different point slab every step"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def c4_model_object():
    """The C4 model as a Python model OBJECT (TesseroidModel with a per-cell density list), the
    form a user of the reference's call surface holds; used for the cold end-to-end number."""
    from tesseroid_variable_density_b200.mesher import TesseroidModel
    rng = np.random.default_rng(SEED)
    thickness = rng.uniform(200, 6000, NLAT * NLON)
    top = np.full(NLAT * NLON, 845.0)
    model = TesseroidModel((-44.99, -25.01, 280.01, 299.99), top, top - thickness, (NLAT, NLON))
    model.addprop("density", [c4_density()] * model.size)
    return model


def irregular_bounds(bounds, seed=7):
    """C4 bounds with every tesseroid's longitude extent made distinct (jitter of 1e-7 deg):
    no two records share (w, e), so the sweep recomputes the longitude half of the quadrature
    for every tesseroid -- what an irregular (non-mesh) model costs."""
    rng = np.random.default_rng(seed)
    b = bounds.copy()
    b[:, 0] += rng.uniform(-1e-7, 1e-7, b.shape[0])
    b[:, 1] += rng.uniform(-1e-7, 1e-7, b.shape[0])
    return b


class Timer(object):
    """CUDA-event timing on torch's current stream, max over ranks."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, n):
        """device ms of `for i in range(n): fn(i)`, max over ranks"""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.current_stream()
        e0.record(stream)
        for i in range(n):
            fn(i)
        e1.record(stream)
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def wall_max(self, seconds):
        t = self.torch.tensor([seconds], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def phase_ms(lib, reset):
    out = (C.c_double * 4)()
    lib.tvd_phase_ms(1 if reset else 0, out)
    return [float(v) for v in out]


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from tesseroid_variable_density_b200 import _lib, tesseroid
    from tesseroid_variable_density_b200.model import FlatModel
    from tesseroid_variable_density_b200.tesseroid import (PreparedModel, _convert_coords,
                                                           forward_raw, STACK_SIZE)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # host-side barrier for the cold leg: ranks waiting in an NCCL barrier would spin on
        # the GPUs that rank 0 is about to use
        cpu_group = dist.new_group(backend="gloo")
    lib = _lib.load()
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    timer = Timer(torch, dist, world, dev)

    flat = c4_flat_model()
    prepared = PreparedModel(flat, delta=0.1, device=local)
    n_tess, n_pieces = prepared.n_tesseroids, prepared.n_pieces
    lons, lats, heights = c4_points()
    P = args.points
    n_total = lons.size
    # rank r owns the r-th contiguous 1/world of the grid and walks it in slabs of P points
    shard = n_total // world
    base = rank * shard

    def slab(step, z=None, jitter=False):
        start = base + (step * P) % max(1, shard - P + 1) if shard > P else base
        idx = (start + np.arange(P)) % n_total
        h = heights[idx] if z is None else np.full(P, z)
        if jitter:      # topography-following points: no two at the same height
            h = h + np.random.default_rng(step).uniform(0.0, 500.0, P)
        return lons[idx], lats[idx], h

    stats = np.zeros(_lib.NSTATS, dtype=np.int64)

    def device_call(model, d_pts, d_res, n, n_groups):
        rc = lib.tvd_forward_device(
            _lib.FIELDS["gz"], model.handle, local, n,
            d_pts[0].data_ptr(), d_pts[1].data_ptr(), d_pts[2].data_ptr(), d_pts[3].data_ptr(),
            2.5, STACK_SIZE, n_groups, d_res.data_ptr(), _lib.ptr_i64(stats),
            C.c_void_p(stream.cuda_stream))
        _lib.check(rc, "tvd_forward_device")

    def to_device(cc):
        return torch.from_numpy(np.stack(cc)).to(dev)

    # ---- headline: weak-scaled slabs at 10 km, inputs resident in HBM ---------------------------
    n_groups = lib.tvd_plan_groups(prepared.handle, P)
    n_steps_total = args.warmup + args.steps
    resident = [to_device(_convert_coords(*slab(s))) for s in range(n_steps_total)]
    d_res = torch.empty(P, dtype=torch.float64, device=dev)
    for s in range(args.warmup):
        device_call(prepared, resident[s], d_res, P, n_groups)
    timer.barrier()
    lib.tvd_launch_count(1)
    phase_ms(lib, True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    agg = np.zeros(_lib.NSTATS, dtype=np.int64)

    def step(i):
        device_call(prepared, resident[args.warmup + i], d_res, P, n_groups)
        agg.__iadd__(stats)
    ms = timer.run(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.tvd_launch_count(0))
    far_ms, _, _, nfar = phase_ms(lib, True)
    checksum = float(d_res.sum().item())
    del resident

    # ---- end to end through the call surface with HOST buffers: _convert_coords on the host,
    # H2D of the step's points and D2H of its result inside the timed region ---------------------
    for s in range(min(2, args.warmup)):
        forward_raw("gz", *_convert_coords(*slab(s)), prepared, 2.5, devices=[local])
    timer.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        forward_raw("gz", *_convert_coords(*slab(args.warmup + s)), prepared, 2.5, devices=[local])
    torch.cuda.synchronize()
    t_e2e = timer.wall_max(time.perf_counter() - t0)

    legs = set(x for x in args.legs.split(",") if x)
    if args.no_cold:
        legs.discard("cold")
    # ---- near-field workload: the same slabs at 2 km (millions of near-field pairs) -------------
    near_steps = max(1, min(3, args.steps)) if "near" in legs else 0
    agg2 = np.zeros(_lib.NSTATS, dtype=np.int64)
    ms2, far2, near2, post2 = 0.0, 0.0, 0.0, 0.0
    if near_steps:
        res2 = [to_device(_convert_coords(*slab(s, z=2e3))) for s in range(near_steps + 1)]
        device_call(prepared, res2[0], d_res, P, n_groups)
        phase_ms(lib, True)

        def step2(i):
            device_call(prepared, res2[1 + i], d_res, P, n_groups)
            agg2.__iadd__(stats)
        ms2 = timer.run(step2, near_steps)
        far2, near2, post2, _ = phase_ms(lib, True)
        del res2

    # ---- bracket of the headline: what an irregular model / topography-following points cost ----
    bracket = {}
    br_steps = max(1, min(2, args.steps))
    irr = PreparedModel(FlatModel(irregular_bounds(flat.bounds), flat.kind, flat.params),
                        delta=0.1, device=local) if "bracket" in legs else None
    for name, model, jitter in ((("irregular_model", irr, False),
                                 ("topography_points", prepared, True),
                                 ("irregular_model_and_topography_points", irr, True))
                                if "bracket" in legs else ()):
        pts = [to_device(_convert_coords(*slab(s, jitter=jitter))) for s in range(br_steps + 1)]
        device_call(model, pts[0], d_res, P, n_groups)
        phase_ms(lib, True)
        t = timer.run(lambda i: device_call(model, pts[1 + i], d_res, P, n_groups), br_steps)
        f_ms = phase_ms(lib, True)[0]
        bracket[name] = {"value": float(n_tess) * P * br_steps * world / (t * 1e-3), "unit": UNIT,
                         "sweep_ms_per_step": f_ms / br_steps}
        del pts
    if irr is not None:
        irr.close()

    ms_strong, identical, g_full, lo, hi, strong_clocks = None, None, None, 0, 0, None
    if "strong" in legs:
        ms_strong, identical, g_full, lo, hi, strong_clocks = strong_leg(
            torch, dist, lib, timer, prepared, device_call, to_device, _convert_coords, lons, lats,
            heights, world, rank, dev)
    # ---- cold end to end: tesseroid.gz(lon, lat, height, model, njobs=N) from a Python model
    # OBJECT to the host result, wall clock, on rank 0 driving all N GPUs of the box in-process
    # (the reference's njobs surface, tesseroid.py:179-195); the other ranks wait ------------------
    timer.barrier()
    del d_res
    torch.cuda.empty_cache()
    e2e_cold = None
    if rank == 0 and "cold" in legs:
        t0 = time.perf_counter()
        model_obj = c4_model_object()
        t_obj = time.perf_counter() - t0
        njobs = min(world, lib.tvd_device_count())
        t0 = time.perf_counter()
        out = tesseroid.gz(lons, lats, heights, model_obj, ratio=2.5, delta=0.1, njobs=njobs)
        t_cold = time.perf_counter() - t0
        e2e_cold = {"seconds": t_cold, "value": float(n_tess) * n_total / t_cold, "unit": UNIT,
                    "call": "tesseroid.gz(lon, lat, height, TesseroidModel(1e6 cells), njobs=%d)"
                            % njobs,
                    "points": int(n_total), "model_object_build_s": t_obj,
                    # CUDA contexts this process had to create inside the call (~0.2 s each,
                    # driver time): the first use of the other GPUs of the box
                    "cuda_contexts_created_inside": njobs - 1,
                    "checksum": float(np.sum(out))}
        if njobs > 1:
            # the same call again: the model is flattened, discretised and replicated from the
            # Python objects once more, but the devices are initialised
            t0 = time.perf_counter()
            out = tesseroid.gz(lons, lats, heights, model_obj, ratio=2.5, delta=0.1, njobs=njobs)
            t_again = time.perf_counter() - t0
            e2e_cold["second_call_seconds"] = t_again
            e2e_cold["second_call_value"] = float(n_tess) * n_total / t_again
    if cpu_group is not None:
        dist.barrier(group=cpu_group)
    timer.barrier()

    if rank == 0:
        evals = float(n_tess) * P * args.steps * world
        value = evals / (ms * 1e-3)
        si = _lib.STAT_NAMES.index
        far_pairs = float(agg[si("nodes")] - agg[si("nonroot_nodes")] - agg[si("near_pairs")])
        achieved = far_pairs * F_PAIR_EXECUTED / (far_ms * 1e-3) / 1e12 if far_ms > 0 else None
        achieved_cal = far_pairs * F_PAIR_CALIBRATED / (far_ms * 1e-3) / 1e12 if far_ms > 0 else None
        achieved_v1 = far_pairs * F_PAIR_V1_ESTIMATE / (far_ms * 1e-3) / 1e12 if far_ms > 0 else None
        peak = float(lib.tvd_measure_fp64_peak(local, 8192))
        near_leaves = float(agg2[si("leaves")] - (agg2[si("nodes")] - agg2[si("nonroot_nodes")]
                                                  - agg2[si("near_pairs")]))
        near_nodes = float(agg2[si("nonroot_nodes")] + agg2[si("near_pairs")])
        strong_value = float(n_tess) * n_total / (ms_strong * 1e-3) if ms_strong else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args),
            "pieces": n_pieces, "pieces_per_tesseroid": n_pieces / n_tess,
            "leaves_per_eval": float(agg[1]) / (float(n_tess) * P * args.steps),
            "nodes_per_eval": float(agg[0]) / (float(n_tess) * P * args.steps),
            "near_pairs_per_step": float(agg[7]) / args.steps,
            "piece_pairs_per_s": float(n_pieces) * P * args.steps * world / (ms * 1e-3),
            "sweep_groups": int(n_groups),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if (achieved and peak > 0) else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this
                         # kernel at the default shape (ncu --set full, profiles/far_sweep_r2.md);
                         # the algorithmic minimum is the piece records read once (449 MB) plus
                         # the group sums written once (78 MB)
                         "traffic": FAR_TRAFFIC_BYTES if P == 75776 else None,
                         "kernel": "far_sweep<gz>",
                         "kernel_ms_per_launch": far_ms / max(1.0, nfar),
                         "kernel_share_of_step": far_ms / ms,
                         "flops_per_piece_pair": F_PAIR_EXECUTED,
                         "piece_pairs_per_launch": far_pairs / max(1.0, nfar),
                         # the instruction mix (DADD/DMUL count 1 flop, DFMA 2) caps the flop
                         # fraction of a saturated FP64 pipe at 128.8 / (2 x 88.4) = 0.73
                         "frac_ceiling_of_this_instruction_mix": 0.73,
                         # measured rate of FP64 INSTRUCTIONS against the bare DFMA loop's (this
                         # run's `peak`): achieved / peak / (128.8 / (2 x 88.4))
                         "fp64_instruction_rate_vs_bare_dfma_loop":
                             (achieved / peak / (F_PAIR_EXECUTED / (2.0 * 88.4)))
                             if (achieved and peak > 0) else None,
                         "achieved_with_first_kernel_count_218": achieved_cal,
                         "achieved_with_survey_estimate_384": achieved_v1,
                         "peak_source": "DFMA microbenchmark run by this process "
                                        "(MEASURED_PEAKS.json has no FP64 figure)",
                         # static facts of this kernel build from its ncu capture; not measured
                         # by this run
                         "ncu": FAR_NCU},
            "e2e": {"value": float(n_tess) * P * args.steps * world / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": 4 * 8 * P, "d2h_bytes_per_step": 8 * P},
            # second timed workload: the same slabs 2 km above the model (near-field pairs)
            "near": None if not near_steps else {
                     "workload": "same slabs at z = 2 km", "steps": near_steps,
                     "value": float(n_tess) * P * near_steps * world / (ms2 * 1e-3), "unit": UNIT,
                     "ms_per_step": ms2 / near_steps,
                     "near_pairs_per_step": float(agg2[si("near_pairs")]) / near_steps,
                     "walk_ms_per_step": near2 / near_steps,
                     "walk_share_of_step": near2 / ms2 if ms2 > 0 else None,
                     "after_sweep_ms_per_step": post2 / near_steps,
                     "sweep_ms_per_step": far2 / near_steps,
                     "near_leaves_per_s": near_leaves / (near2 * 1e-3) if near2 > 0 else None,
                     "near_nodes_per_s": near_nodes / (near2 * 1e-3) if near2 > 0 else None,
                     "slowdown_vs_10km": (ms2 / near_steps) / (ms / args.steps)},
            # the headline rate assumes a regular mesh and points at one height
            "bracket": bracket,
            # fixed-size scaling: the whole 1e6-point grid split over the ranks, one pass
            "strong": None if not ms_strong else {
                       "value": strong_value, "unit": UNIT, "seconds": ms_strong * 1e-3,
                       "points_per_gpu": int(hi - lo), "sweep_groups": int(g_full),
                       "efficiency_vs_weak_rate": strong_value / value,
                       "sharding_bit_identical": identical, "clocks": strong_clocks},
            "sharding_bit_identical": identical,
            "e2e_cold": e2e_cold,
            "gpu_launches": launches, "clocks": clocks, "checksum": checksum,
        }
        if world == 1 and not args.no_cpu:
            # separate process: the reference's njobs path forks a multiprocessing pool, which
            # must not inherit this process's CUDA context
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl",
                                  "cpu-sample"], capture_output=True, text=True)
            try:
                line["cpu_baseline"] = json.loads(out.stdout.strip().splitlines()[-1])
            except Exception:
                line["cpu_baseline"] = {"error": (out.stderr or out.stdout)[-300:]}
        print(json.dumps(line), flush=True)
    prepared.close()
    if world > 1:
        dist.destroy_process_group()


def strong_leg(torch, dist, lib, timer, prepared, device_call, to_device, _convert_coords, lons,
               lats, heights, world, rank, dev):
    """Fixed-size (strong) scaling: the WHOLE 1000 x 1000 grid split over the ranks, one pass;
    then the sharding check.  Returns (ms, bit_identical, n_groups, lo, hi, clocks)."""
    n_total = lons.size
    # contiguous chunks as _split_arrays (tesseroid.py:318-323), global summation plan
    from tesseroid_variable_density_b200.distributed import shard_bounds
    g_full = lib.tvd_plan_groups(prepared.handle, n_total)
    lo, hi = shard_bounds(n_total, world, rank)
    full_pts = to_device(_convert_coords(lons[lo:hi], lats[lo:hi], heights[lo:hi]))
    full_res = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    device_call(prepared, resident_small(full_pts, 4096), full_res[:4096], 4096, g_full)   # warm-up
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    ms_strong = timer.run(lambda i: device_call(prepared, full_pts, full_res, hi - lo, g_full), 1)
    strong_clocks = sampler.stop() if rank == 0 else None
    # sharding check: rank 0 recomputes the first 4096 points of every rank's chunk (and, with
    # one rank, a chunk that starts off a multiple of 32) and compares the bits
    probe = 4096
    mine = full_res[:probe].clone()
    if world > 1:
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
    identical = True
    if rank == 0:
        checks = [(r, shard_bounds(n_total, world, r)[0]) for r in range(world)]
        if world == 1:
            checks = [(0, 0), (0, 100003)]
        for r, start in checks:
            pts = to_device(_convert_coords(lons[start:start + probe], lats[start:start + probe],
                                            heights[start:start + probe]))
            out = torch.empty(probe, dtype=torch.float64, device=dev)
            device_call(prepared, pts, out, probe, g_full)
            want = gathered[r] if world > 1 else full_res[start:start + probe]
            identical = identical and bool(torch.equal(out, want.to(dev)))
    del full_pts

    del full_res
    return ms_strong, identical, g_full, lo, hi, strong_clocks


def resident_small(pts, n):
    """the first n points of a resident [4][P] tensor as a contiguous [4][n] tensor"""
    return pts[:, :n].contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=75776,
                    help="points per GPU per step (default 148 SMs x 4 blocks x 128 threads)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cpu-sample"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cold", action="store_true", help="skip the cold end-to-end leg")
    ap.add_argument("--legs", default="near,bracket,strong,cold",
                    help="comma-separated extra legs after the headline and e2e measurements "
                         "(profiling runs use fewer)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "cpu-sample":
        print(json.dumps(cpu_baseline()), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
