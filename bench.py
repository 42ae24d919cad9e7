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
#    capture committed with it (profiles/far_sweep_r1g.md: 11.4 DADD + 44.7 DMUL + 40.3 DFMA =
#    136.7 flops; smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on x elapsed cycles /
#    pairs).  `roofline.achieved` uses this figure: it is the hardware flop rate of the kernel
#    and can be checked against ncu directly.  It must be re-measured when the kernel changes.
#  * SURVEY.md section 8d (convention v1) ESTIMATED 384 = F_test 76 + F_glq 308 from guessed
#    special-function weights and asked for a one-time ncu calibration on the first kernel; that
#    kernel (profiles/far_sweep_r1a.md) executed 218 flops per pair.
#    Later kernels do the same pairs with fewer operations (cheaper test and cos, angular work
#    shared between pieces), so rates computed with 218 or 384 exceed what the hardware executes;
#    they are reported beside `achieved`, not as `achieved`.
F_PAIR_EXECUTED = 136.7
F_PAIR_CALIBRATED = 218
F_PAIR_V1_ESTIMATE = 384


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
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from tesseroid_variable_density_b200 import _lib
    from tesseroid_variable_density_b200.tesseroid import (PreparedModel, _convert_coords,
                                                           forward_raw, STACK_SIZE)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    flat = c4_flat_model()
    prepared = PreparedModel(flat, delta=0.1, device=local)
    n_tess, n_pieces = prepared.n_tesseroids, prepared.n_pieces
    lons, lats, heights = c4_points()
    P = args.points
    n_total = lons.size
    # rank r owns the r-th contiguous 1/world of the grid and walks it in slabs of P points
    shard = n_total // world
    base = rank * shard

    def slab(step):
        start = base + (step * P) % max(1, shard - P + 1) if shard > P else base
        idx = (start + np.arange(P)) % n_total
        return lons[idx], lats[idx], heights[idx]

    n_steps_total = args.warmup + args.steps
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    stats = np.zeros(_lib.NSTATS, dtype=np.int64)
    n_groups = lib.tvd_plan_groups(prepared.handle, P)

    def device_step(d_pts, d_res):
        rc = lib.tvd_forward_device(
            _lib.FIELDS["gz"], prepared.handle, local, P,
            d_pts[0].data_ptr(), d_pts[1].data_ptr(), d_pts[2].data_ptr(), d_pts[3].data_ptr(),
            2.5, STACK_SIZE, n_groups, d_res.data_ptr(), _lib.ptr_i64(stats),
            C.c_void_p(stream.cuda_stream))
        _lib.check(rc, "tvd_forward_device")

    # inputs resident in HBM before the timed region
    resident = []
    for s in range(n_steps_total):
        cc = _convert_coords(*slab(s))
        resident.append(torch.from_numpy(np.stack(cc)).to(dev))
    d_res = torch.empty(P, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        device_step(resident[s], d_res)
    barrier()
    lib.tvd_launch_count(1)
    lib.tvd_far_kernel_ms(1, None)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = np.zeros(_lib.NSTATS, dtype=np.int64)
    e0.record(stream)
    for s in range(args.steps):
        device_step(resident[args.warmup + s], d_res)
        agg += stats
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.tvd_launch_count(0))
    nfar = C.c_int64(0)
    far_ms = float(lib.tvd_far_kernel_ms(0, C.byref(nfar)))
    checksum = float(d_res.sum().item())

    # end to end through the public call surface with HOST buffers: _convert_coords on the
    # host, H2D of the step's points and D2H of its result inside the timed region
    for s in range(min(2, args.warmup)):
        lo, la, he = slab(s)
        forward_raw("gz", *_convert_coords(lo, la, he), prepared, 2.5, devices=[local])
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo, la, he = slab(args.warmup + s)
        out = forward_raw("gz", *_convert_coords(lo, la, he), prepared, 2.5, devices=[local])
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())

    if rank == 0:
        evals = float(n_tess) * P * args.steps * world
        value = evals / (ms * 1e-3)
        # roofline of the dominant kernel (far-field sweep): algorithmic FP64 flops by the
        # canonical convention divided by its CUDA-event duration on its launch stream
        far_pairs = float(agg[_lib.STAT_NAMES.index("nodes")] - agg[_lib.STAT_NAMES.index("nonroot_nodes")]
                          - agg[_lib.STAT_NAMES.index("near_pairs")])
        achieved = far_pairs * F_PAIR_EXECUTED / (far_ms * 1e-3) / 1e12 if far_ms > 0 else None
        achieved_cal = far_pairs * F_PAIR_CALIBRATED / (far_ms * 1e-3) / 1e12 if far_ms > 0 else None
        achieved_v1 = far_pairs * F_PAIR_V1_ESTIMATE / (far_ms * 1e-3) / 1e12 if far_ms > 0 else None
        peak = float(lib.tvd_measure_fp64_peak(local, 8192))
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
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if (achieved and peak > 0) else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this
                         # kernel at the default shape (ncu --set full, profiles/far_sweep_r1g.md);
                         # the algorithmic minimum is the piece records read once, 337 MB
                         "traffic": 370.1e6 if P == 75776 else None,
                         "kernel": "far_sweep<gz>",
                         "kernel_ms_per_launch": far_ms / max(1, nfar.value),
                         "kernel_share_of_step": far_ms / ms,
                         "flops_per_piece_pair": F_PAIR_EXECUTED,
                         "piece_pairs_per_launch": far_pairs / max(1, nfar.value),
                         # the instruction mix (DADD/DMUL count 1 flop, DFMA 2) caps the flop
                         # fraction of a saturated FP64 pipe at 136.7 / (2 x 96.4) = 0.71
                         "frac_ceiling_of_this_instruction_mix": 0.71,
                         "achieved_with_first_kernel_count_218": achieved_cal,
                         "achieved_with_survey_estimate_384": achieved_v1,
                         "peak_source": "DFMA microbenchmark run by this process "
                                        "(MEASURED_PEAKS.json has no FP64 figure)",
                         # static facts of this kernel build from its ncu capture
                         # (profiles/far_sweep_r1g.md); not measured by this run
                         "ncu": {"fp64_pipe_pct_of_peak": 86.8, "issue_active_pct": 59.9,
                                 "executed_fp64_instructions_per_piece_pair": 96.4,
                                 "executed_flops_per_piece_pair": 136.7}},
            "e2e": {"value": float(n_tess) * P * args.steps * world / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": 4 * 8 * P, "d2h_bytes_per_step": 8 * P},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=75776,
                    help="points per GPU per step (default 148 SMs x 4 blocks x 128 threads)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cpu-sample"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
