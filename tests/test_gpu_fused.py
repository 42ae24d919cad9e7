"""
Fused multi-field sweeps against the single-field calls: identical counters, and
values equal to a few ulp (bit-identical when the set uses the same l^-3 evaluation as the single
fields: {gx, gy, gz} and the six tensor components).
"""
import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import gridder, tesseroid
from tesseroid_variable_density_b200.tesseroid import (PreparedModel, _convert_coords, forward_raw,
                                                       forward_raw_fused, fused_supported)

pytestmark = pytest.mark.gpu

SETS = [
    ["potential", "gz"],
    ["potential", "gz", "gzz"],
    ["potential", "gx", "gy", "gz"],
    ["gx", "gy", "gz"],
    ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz"],
    ["gzz", "potential", "gz"],           # caller order differs from slot order
]


@pytest.mark.parametrize("names", SETS, ids=["-".join(s) for s in SETS])
@pytest.mark.parametrize("grid", ["global2km", "260km"])
def test_fused_equals_single_field_calls(names, grid):
    mesh = cases.sine_case(1e5, 2)[0]
    lats, lons, heights = cases.GRIDS[grid]()
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    ratios = [cases.DEFAULT_RATIO[f] for f in names]
    assert fused_supported(names)
    fused, fstats = forward_raw_fused(names, *cc, pm, ratios, return_stats=True)
    for f, r in zip(names, ratios):
        single, sstats = forward_raw(f, *cc, pm, r, return_stats=True)
        if set(names) <= {"gx", "gy", "gz"} or set(names) <= set(cases.ALL_FIELDS[4:]):
            assert np.array_equal(fused[f], single), f
        else:
            scale = max(np.max(np.abs(forward_raw(g, *cc, pm, cases.DEFAULT_RATIO[g])))
                        for g in names if cases.UNIT[g] == cases.UNIT[f])
            assert np.max(np.abs(fused[f] - single)) <= 1e-14 * scale, f
        assert fstats[f] == sstats, f
    pm.close()


def test_fused_regional_with_near_field(oracle_lib):
    from test_gpu_parity import regional_model
    from tesseroid_variable_density_b200.model import flatten_model
    tm = regional_model(nlat=20, nlon=24)
    lats, lons, heights = gridder.regular((-39.9, -39.1, 288.1, 289.1), (25, 25), z=2000.)
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(tm, delta=0.1)
    names = ["potential", "gz", "gzz"]
    fused, st = forward_raw_fused(names, *cc, pm, [1, 2.5, 8], return_stats=True)
    pm.close()
    assert st["gzz"]["near_pairs"] > st["gz"]["near_pairs"] > st["potential"]["near_pairs"] > 0
    flat = flatten_model(tm)
    for f, r, tol in (("potential", 1, 1e-9), ("gz", 2.5, 1e-9), ("gzz", 8, 1e-6)):
        o, so = oracle_lib.forward(f, flat.bounds, flat.kind, flat.params, 0.1, r, *cc, n_threads=8)
        assert np.max(np.abs(fused[f] - o)) <= tol * np.max(np.abs(o)), f
        assert all(st[f][k] == so[k] for k in ("nodes", "leaves", "splits")), f


def test_fields_api():
    mesh, _, exact = cases.exponential_case(1e4, 10)
    lats, lons, heights = cases.GRIDS["260km"]()
    out = tesseroid.fields(["potential", "gz", "gzz"], lons, lats, heights, mesh, delta=0.1)
    for f in ("potential", "gz", "gzz"):
        single = getattr(tesseroid, f)(lons, lats, heights, mesh, delta=0.1)
        assert np.max(np.abs(out[f] - single)) <= 1e-14 * np.max(np.abs(single))
    ex = exact(heights[0])
    for f in out:
        assert cases.percent_difference(out[f], ex[f]) <= 0.1
    # a set that is not compiled in falls back to field-by-field evaluation
    assert not fused_supported(["gx", "gzz"])
    out = tesseroid.fields(["gx", "gzz"], lons, lats, heights, mesh, delta=0.1)
    assert np.array_equal(out["gzz"], tesseroid.gzz(lons, lats, heights, mesh, delta=0.1))


def test_fused_device_entry_point():
    """tvd_forward_fused_device (device-resident points and results) = the host entry point."""
    import ctypes as C
    import torch
    from tesseroid_variable_density_b200 import _lib
    lib = _lib.load()
    mesh = cases.exponential_case(1e4, 10)[0]
    lats, lons, heights = cases.GRIDS["global2km"]()
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    names, ratios = ["gzz", "potential", "gz"], [8.0, 1.0, 2.5]
    want = forward_raw_fused(names, *cc, pm, ratios)
    n = lons.size
    d = [torch.from_numpy(np.ascontiguousarray(c)).cuda() for c in cc]
    out = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in names]
    ids = (C.c_int * 3)(*[_lib.FIELDS[f] for f in names])
    rat = np.ascontiguousarray(ratios, dtype=np.float64)
    ptrs = (C.c_void_p * 3)(*[o.data_ptr() for o in out])
    stats = np.zeros((3, _lib.NSTATS), dtype=np.int64)
    rc = lib.tvd_forward_fused_device(3, ids, _lib.ptr_f64(rat), pm.handle, 0, n, d[0].data_ptr(),
                                      d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), 200,
                                      lib.tvd_plan_groups(pm.handle, n), ptrs, _lib.ptr_i64(stats),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    for f, o in zip(names, out):
        assert np.array_equal(o.cpu().numpy(), want[f]), f
    pm.close()


def test_fused_sweep_is_bit_exact_under_point_sharding():
    """The fused sweep sharded over chunks (the njobs path, here three chunks on device 0) returns
    the bits of the unsharded call for every field of the set."""
    mesh = cases.sine_case(1e5, 2)[0]
    lats, lons, heights = cases.GRIDS["global2km"]()
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    names, ratios = ["potential", "gz", "gzz"], [1, 2.5, 8]
    one, s1 = forward_raw_fused(names, *cc, pm, ratios, return_stats=True)
    many, s3 = forward_raw_fused(names, *cc, pm, ratios, devices=[0, 0, 0], return_stats=True)
    pm.close()
    for f in names:
        assert np.array_equal(one[f], many[f]), f
        assert s1[f] == s3[f], f
