"""Size-independent properties of the hot path at sizes the oracle cannot finish."""
import numpy as np
import pytest

from tesseroid_variable_density_b200 import density as D
from tesseroid_variable_density_b200 import gridder, _lib
from tesseroid_variable_density_b200.model import FlatModel
from tesseroid_variable_density_b200.tesseroid import (PreparedModel, _convert_coords, forward_raw,
                                                      forward_raw_fused)

pytestmark = pytest.mark.gpu


def synthetic(nlat, nlon, seed=20190277):
    """A patch of the C4 synthetic regional model (bench.py) with nlat x nlon cells of 0.02 deg."""
    rng = np.random.default_rng(seed)
    thickness = rng.uniform(200, 6000, nlat * nlon)
    j = np.tile(np.arange(nlon), nlat)
    i = np.repeat(np.arange(nlat), nlon)
    b = np.empty((nlat * nlon, 6))
    b[:, 0] = 280.0 + 0.02 * j
    b[:, 1] = b[:, 0] + 0.02
    b[:, 2] = -45.0 + 0.02 * i
    b[:, 3] = b[:, 2] + 0.02
    b[:, 4] = 845.0
    b[:, 5] = 845.0 - thickness
    return b


def flat_with(bounds, dens):
    n = bounds.shape[0]
    k, p = (dens.kind, dens.params) if hasattr(dens, "kind") else (0, (dens, 0, 0, 0, 0))
    return FlatModel(bounds, np.full(n, k, dtype=np.int32), np.tile(p, (n, 1)))


@pytest.fixture(scope="module")
def big():
    bounds = synthetic(300, 300)                      # 90 000 tesseroids
    lats, lons, heights = gridder.regular((-44.9, -39.1, 280.1, 285.9), (200, 200), z=2e3)
    return bounds, _convert_coords(lons, lats, heights)


def test_linearity_in_density(big):
    bounds, cc = big
    a = D.ExponentialDensity(-150., -3., -5155., 6000., -270.)
    a2 = D.ExponentialDensity(-300., -3., -5155., 6000., -540.)     # exactly 2x (powers of two)
    ra = forward_raw("gz", *cc, PreparedModel(flat_with(bounds, a), delta=0.1), 2.5)
    r2 = forward_raw("gz", *cc, PreparedModel(flat_with(bounds, a2), delta=0.1), 2.5)
    assert np.array_equal(r2, 2 * ra)       # scaling the density by 2 is exact in binary FP


def test_additivity_over_model_halves(big):
    bounds, cc = big
    dens = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
    whole = forward_raw("gz", *cc, PreparedModel(flat_with(bounds, dens), delta=0.1), 2.5)
    h = bounds.shape[0] // 2
    p1 = forward_raw("gz", *cc, PreparedModel(flat_with(bounds[:h], dens), delta=0.1), 2.5)
    p2 = forward_raw("gz", *cc, PreparedModel(flat_with(bounds[h:], dens), delta=0.1), 2.5)
    assert np.max(np.abs(whole - (p1 + p2))) <= 1e-12 * np.max(np.abs(whole))


def test_point_sharding_is_bit_exact(big):
    """Any contiguous chunk of the points, evaluated alone with the same summation plan, gives
    the bits of the full run (the reference's njobs property, tesseroid.py:188-192)."""
    import ctypes as C
    bounds, cc = big
    dens = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
    pm = PreparedModel(flat_with(bounds, dens), delta=0.1)
    full, stats = forward_raw("gz", *cc, pm, 2.5, return_stats=True)
    assert stats["near_pairs"] > 0              # the near-field path is exercised
    n = full.size
    lib = _lib.load()
    groups = lib.tvd_plan_groups(pm.handle, n)
    import torch
    for lo, hi in ((0, 1), (17, 5000), (n // 3, n // 3 + 12345), (n - 777, n)):
        d = [torch.from_numpy(np.ascontiguousarray(c[lo:hi])).cuda() for c in cc]
        out = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
        st = np.zeros(8, dtype=np.int64)
        rc = lib.tvd_forward_device(3, pm.handle, 0, hi - lo, d[0].data_ptr(), d[1].data_ptr(),
                                    d[2].data_ptr(), d[3].data_ptr(), 2.5, 200, groups,
                                    out.data_ptr(), _lib.ptr_i64(st),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        assert np.array_equal(out.cpu().numpy(), full[lo:hi])
    pm.close()


def test_repeatability_and_stats_accounting(big):
    bounds, cc = big
    dens = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
    pm = PreparedModel(flat_with(bounds, dens), delta=0.1)
    r1, s1 = forward_raw("gz", *cc, pm, 2.5, return_stats=True)
    r2, s2 = forward_raw("gz", *cc, pm, 2.5, return_stats=True)
    assert np.array_equal(r1, r2) and s1 == s2      # atomics in the queue do not leak into values
    P, Q = cc[0].size, pm.n_pieces
    assert s1["pieces"] == Q
    assert s1["nodes"] == P * Q + s1["nonroot_nodes"]
    assert s1["leaves"] == s1["nodes"] - s1["splits"]
    pm.close()


def test_shell_symmetry_at_scale():
    """A fine shell (6 480 tesseroids) seen from 260 km: gz constant, gx = gy = 0 to the
    quadrature error."""
    from tesseroid_variable_density_b200 import tesseroid, analytical
    from tesseroid_variable_density_b200.mesher import TesseroidMesh
    mesh = TesseroidMesh((0, 360, -90, 90, 0, -3e4), (1, 54, 120))
    dens = D.ExponentialDensity.between(2670, 3300, 0., -3e4, 10)
    mesh.addprop("density", [dens] * mesh.size)
    lats, lons, heights = gridder.regular((-90, 90, 0, 360), (61, 91), z=260e3)
    pm = PreparedModel(mesh, delta=0.1)
    gz = tesseroid.gz(lons, lats, heights, pm)
    gx = tesseroid.gx(lons, lats, heights, pm)
    pm.close()
    A = (3300 - 2670) / (1 - np.exp(-10))
    exact = analytical.shell_exponential_density(260e3, 0., -3e4, A, 10, 3300 - A)["gz"]
    assert np.max(np.abs(gz - exact)) <= 1e-3 * exact
    assert np.max(np.abs(gx)) <= 1e-3 * exact


def test_queue_overflow_is_refilled_not_recomputed(big, monkeypatch):
    """A near-field queue that is too small is rebuilt by a test-only sweep; values, counters
    and the deferred pairs are those of a run with ample capacity."""
    bounds, cc = big
    dens = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
    flat = flat_with(bounds, dens)
    pm = PreparedModel(flat, delta=0.1)
    want, swant = forward_raw("gz", *cc, pm, 2.5, return_stats=True)
    pm.close()
    assert swant["near_pairs"] > 1000
    monkeypatch.setenv("TVD_DEBUG_QCAP", "1000")
    pm = PreparedModel(flat, delta=0.1)
    got, sgot = forward_raw("gz", *cc, pm, 2.5, return_stats=True)
    pm.close()
    assert np.array_equal(got, want) and sgot == swant


def test_one_height_specialisation_is_bit_exact(big):
    """Points at one height take a sweep with 2r*rc and r^2 + rc^2 precomputed per piece; adding
    one point at another height switches the specialisation off.  Same bits either way."""
    bounds, cc = big
    dens = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
    pm = PreparedModel(flat_with(bounds, dens), delta=0.1)
    n = 20000
    lon, sinlat, coslat, radius = (c[:n].copy() for c in cc)
    assert np.all(radius == radius[0])
    uniform = forward_raw("gz", lon, sinlat, coslat, radius, pm, 2.5)
    groups = _lib.load().tvd_plan_groups(pm.handle, n)
    assert groups == _lib.load().tvd_plan_groups(pm.handle, n + 1)     # same summation plan
    mixed = forward_raw("gz", np.append(lon, lon[0]), np.append(sinlat, sinlat[0]),
                        np.append(coslat, coslat[0]), np.append(radius, radius[0] + 123.0), pm, 2.5)
    assert np.array_equal(mixed[:n], uniform)
    pm.close()


def test_root_kernel_expansion_gives_the_walks_bits(big, monkeypatch):
    """Barely-near pairs (root splits once, children are leaves) are finished by the root kernel
    when the queue is large and by the warp walk otherwise; both must return the same bits and
    counters (the root kernel sums the <= 4 leaves in the walk's butterfly order)."""
    bounds, cc = big
    dens = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("TVD_DEBUG_SHALLOW", flag)
        pm = PreparedModel(flat_with(bounds, dens), delta=0.1)
        out[flag] = forward_raw("gz", *cc, pm, 2.5, return_stats=True)
        pm.close()
    assert out["0"][1]["near_pairs"] > 10000
    assert out["0"][1] == out["1"][1]
    assert np.array_equal(out["0"][0], out["1"][0])


def test_point_slabs_do_not_change_a_bit(big, monkeypatch):
    """A device's share of the points is processed in slabs that bound the point-proportional
    device buffers (tvd_abi.cu: point_slab), the way the reference streams points one by one
    (_tesseroid.pyx:65).  The summation plan is that of the whole job, so values, counters and
    per-(tesseroid, point) leaf counts of a slabbed run are those of the single pass - also when
    the slabs are ragged (1000 is not a multiple of the 512-point blocks) and on two chunks."""
    bounds, cc = big
    dens = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
    flat = flat_with(bounds[1500:1900], dens)          # rows under the first rows of points
    n = 5000
    pts = [c[:n].copy() for c in cc]
    pm = PreparedModel(flat, delta=0.1)
    want, swant, cwant = forward_raw("gz", *pts, pm, 2.5, leaf_counts=True)
    assert swant["near_pairs"] > 100
    fwant = forward_raw_fused(("potential", "gz", "gzz"), *pts, pm, (1.0, 2.5, 8.0))
    monkeypatch.setenv("TVD_DEBUG_POINT_SLAB", "1000")
    got, sgot, cgot = forward_raw("gz", *pts, pm, 2.5, leaf_counts=True)
    assert np.array_equal(got, want) and sgot == swant and np.array_equal(cgot, cwant)
    two = forward_raw("gz", *pts, pm, 2.5, devices=[0, 0])
    assert np.array_equal(two, want)
    fgot = forward_raw_fused(("potential", "gz", "gzz"), *pts, pm, (1.0, 2.5, 8.0))
    for name in fwant:
        assert np.array_equal(fgot[name], fwant[name])
    pm.close()
