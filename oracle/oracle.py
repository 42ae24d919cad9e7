"""
ctypes wrapper of ``liboracle.so`` (oracle/tess_oracle.c), the CPU restatement of the
reference hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
NPAR = 5
NSTATS = 8
STAT_NAMES = ("nodes", "leaves", "splits", "too_small", "pieces", "nonroot_nodes",
              "nonroot_leaves", "near_pairs")
FIELDS = {"potential": 0, "gx": 1, "gy": 2, "gz": 3, "gxx": 4, "gxy": 5, "gxz": 6, "gyy": 7,
          "gyz": 8, "gzz": 9}

_lib = None
_dp = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(HERE, "tess_oracle.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.orc_density_eval.restype = C.c_double
        lib.orc_density_eval.argtypes = [C.c_int, _dp, C.c_double]
        lib.orc_density_discretize.restype = C.c_int
        lib.orc_density_discretize.argtypes = [C.c_double, C.c_double, C.c_int, _dp, C.c_double,
                                               C.c_int, _dp, _dp]
        lib.orc_forward.restype = C.c_int
        lib.orc_forward.argtypes = [C.c_int, C.c_int64, _dp, C.POINTER(C.c_int32), _dp,
                                    C.c_double, C.c_double, C.c_int, C.c_int64, _dp, _dp, _dp,
                                    _dp, _dp, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_int]
        _lib = lib
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def density_discretize(top, bottom, kind, params, delta, max_out=4096):
    """(tops, bottoms) of the radial pieces, reference emission order (tesseroid.py:237-283)."""
    lib = load()
    p = _f64(params)
    tops, bots = np.empty(max_out), np.empty(max_out)
    n = lib.orc_density_discretize(float(top), float(bottom), int(kind), p.ctypes.data_as(_dp),
                                   float(delta), max_out, tops.ctypes.data_as(_dp),
                                   bots.ctypes.data_as(_dp))
    if n < 0:
        raise RuntimeError("too many pieces")
    return tops[:n].copy(), bots[:n].copy()


def forward(field, bounds, kind, params, delta, ratio, lon_rad, sinlat, coslat, radius,
            stack_size=200, n_threads=1, leaf_counts=False):
    """
    Field of the whole model in kernel units (field / G), as accumulated by the reference's
    ``_forward_model`` (tesseroid.py:198-234).  Returns ``(result, stats[, counts])``.
    ``delta=None`` disables the density-based discretisation.
    """
    lib = load()
    bounds = _f64(bounds).reshape(-1, 6)
    kind = np.ascontiguousarray(kind, dtype=np.int32)
    params = _f64(params).reshape(-1, NPAR)
    lon_rad, sinlat, coslat, radius = _f64(lon_rad), _f64(sinlat), _f64(coslat), _f64(radius)
    n = lon_rad.size
    result = np.zeros(n)
    stats = np.zeros(NSTATS, dtype=np.int64)
    counts = np.zeros((bounds.shape[0], n), dtype=np.int32) if leaf_counts else None
    rc = lib.orc_forward(
        FIELDS[field], bounds.shape[0], bounds.ctypes.data_as(_dp),
        kind.ctypes.data_as(C.POINTER(C.c_int32)), params.ctypes.data_as(_dp),
        float("nan") if delta is None else float(delta), float(ratio), int(stack_size), n,
        lon_rad.ctypes.data_as(_dp), sinlat.ctypes.data_as(_dp), coslat.ctypes.data_as(_dp),
        radius.ctypes.data_as(_dp), result.ctypes.data_as(_dp),
        counts.ctypes.data_as(C.POINTER(C.c_int32)) if leaf_counts else None,
        stats.ctypes.data_as(C.POINTER(C.c_int64)), int(n_threads))
    if rc == -1:
        raise ValueError("Tesseroid stack overflow")
    if rc != 0:
        raise RuntimeError("orc_forward failed with %d" % rc)
    st = dict(zip(STAT_NAMES, (int(s) for s in stats)))
    if leaf_counts:
        return result, st, counts
    return result, st
