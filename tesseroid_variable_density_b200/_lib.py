"""
ctypes binding of ``libtvd.so`` (C ABI declared in ``include/tvd.h``).

The CUDA library is the only compute path: if it is missing, or no CUDA device is present,
every call raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TVD_LIB_PATH: development knob to load an alternative build of the same library
LIB_PATH = os.environ.get("TVD_LIB_PATH") or os.path.join(_HERE, "csrc", "libtvd.so")

FIELDS = {"potential": 0, "gx": 1, "gy": 2, "gz": 3, "gxx": 4, "gxy": 5, "gxz": 6, "gyy": 7,
          "gyz": 8, "gzz": 9}
NSTATS = 8
STAT_NAMES = ("nodes", "leaves", "splits", "too_small", "pieces", "nonroot_nodes",
              "nonroot_leaves", "near_pairs")
ERR_STACK_OVERFLOW = -1

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol include/tvd.h declares
SIGNATURES = {
    "tvd_device_count": (C.c_int, []),
    "tvd_last_error": (C.c_char_p, []),
    "tvd_warm_devices": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "tvd_model_create": (C.c_int, [C.c_int64, _dp, _ip, _dp, C.c_double, C.c_int,
                                   C.POINTER(C.c_void_p)]),
    "tvd_model_free": (None, [C.c_void_p]),
    "tvd_model_num_pieces": (C.c_int64, [C.c_void_p]),
    "tvd_model_get_pieces": (C.c_int, [C.c_void_p, _dp, _dp, _ip]),
    "tvd_forward": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, C.c_double,
                              C.c_int, C.c_int, C.POINTER(C.c_int), _dp, _lp, _ip]),
    "tvd_forward_fused": (C.c_int, [C.c_int, C.POINTER(C.c_int), _dp, C.c_void_p, C.c_int64, _dp,
                                    _dp, _dp, _dp, C.c_int, C.c_int, C.POINTER(C.c_int), _dp, _lp]),
    "tvd_sensitivity": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, C.c_double,
                                  C.c_int, C.c_int, _dp, _lp]),
    "tvd_forward_fused_device": (C.c_int, [C.c_int, C.POINTER(C.c_int), _dp, C.c_void_p, C.c_int,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                           _lp, C.c_void_p]),
    "tvd_fused_supported": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "tvd_forward_device": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                     C.c_int, C.c_void_p, _lp, C.c_void_p]),
    "tvd_plan_groups": (C.c_int, [C.c_void_p, C.c_int64]),
    "tvd_launch_count": (C.c_int64, [C.c_int]),
    "tvd_far_kernel_ms": (C.c_double, [C.c_int, _lp]),
    "tvd_phase_ms": (None, [C.c_int, _dp]),
    "tvd_measure_fp64_peak": (C.c_double, [C.c_int, C.c_int]),
    "tvd_debug_trig": (C.c_int, [C.c_int, C.c_int64, _dp, _dp, _dp]),
    "tvd_debug_math": (C.c_int, [C.c_int, C.c_int, C.c_int64, _dp, _dp]),
}


def load():
    """Load libtvd.so (once) and declare its signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "{} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or tesseroid_variable_density_b200/csrc/build.sh. The CUDA library is the only "
            "compute path (no CPU fallback).".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    msg = load().tvd_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    """Translate a libtvd return code into the exception the reference would raise."""
    if rc == 0:
        return
    msg = last_error()
    if rc == ERR_STACK_OVERFLOW:
        raise ValueError("Tesseroid stack overflow")      # _tesseroid.pyx:88-89
    raise RuntimeError("{} failed with code {}: {}".format(what, rc, msg))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr_f64(a):
    return a.ctypes.data_as(_dp)


def ptr_i32(a):
    return a.ctypes.data_as(_ip)


def ptr_i64(a):
    return a.ctypes.data_as(_lp)
