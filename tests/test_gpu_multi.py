"""1 GPU vs N GPUs: bit identity of the point-sharded run (skipped with a single device)."""
import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import _lib, tesseroid

pytestmark = pytest.mark.gpu


def test_multi_device_bit_identity():
    n = _lib.load().tvd_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    mesh, _, _ = cases.exponential_case(1e4, 10)
    lats, lons, heights = cases.GRIDS["global2km"]()
    one = tesseroid.gz(lons, lats, heights, mesh, delta=0.1, njobs=1)
    for k in sorted({2, min(n, 4), n}):
        many = tesseroid.gz(lons, lats, heights, mesh, delta=0.1, njobs=k)
        assert np.array_equal(one, many)


def test_same_device_twice_equals_sharding():
    """Sharding logic of tvd_forward exercised on one GPU: two chunks on device 0."""
    mesh, _, _ = cases.sine_case(1e5, 2)
    lats, lons, heights = cases.GRIDS["global2km"]()
    one = tesseroid.gz(lons, lats, heights, mesh, delta=0.1)
    two = tesseroid.gz(lons, lats, heights, mesh, delta=0.1, devices=[0, 0, 0])
    assert np.array_equal(one, two)


def test_rank_compute_callable_uses_the_global_plan():
    """distributed.make_gpu_compute: chunks evaluated the way torchrun ranks do (device-resident,
    global summation plan) concatenate to the single-call result, bit for bit."""
    from tesseroid_variable_density_b200.distributed import make_gpu_compute, shard_bounds
    from tesseroid_variable_density_b200.tesseroid import PreparedModel, _convert_coords, forward_raw
    mesh, _, _ = cases.exponential_case(1e4, 10)
    lats, lons, heights = cases.GRIDS["global2km"]()
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    want = forward_raw("gz", *cc, pm, 2.5)
    compute = make_gpu_compute(pm, 2.5, lons.size, device=0)
    got = np.empty_like(want)
    for rank in range(3):
        s, e = shard_bounds(lons.size, 3, rank)
        got[s:e] = compute("gz", *[c[s:e] for c in cc])
    pm.close()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("field", ["gy", "gyy", "gxy", "gz", "potential"])
def test_chunking_is_bit_exact_when_warps_straddle_pi_over_4(field):
    """A point's value must not depend on which points share its warp.  The far-field sweep
    picks its sin/cos code path by warp vote (tvd_far.cu: trig2): on a grid wider than pi/4 in
    longitude, chunk starts that are not multiples of 32 change the composition of every warp
    (ADVICE round 1).  Fields that use sin(lon' - lon) (gy, gxy, gyy, gyz) are the sensitive ones;
    the reference's njobs chunking is bit-exact (tesseroid.py:188-192)."""
    from tesseroid_variable_density_b200 import density as D
    from tesseroid_variable_density_b200 import gridder
    from tesseroid_variable_density_b200.mesher import TesseroidMesh
    from tesseroid_variable_density_b200.tesseroid import PreparedModel, _convert_coords, forward_raw
    mesh = TesseroidMesh((-20, 20, -15, 15, 0, -30000.), (1, 12, 16))
    mesh.addprop("density", [D.ExponentialDensity.between(2670, 3300, 0., -30000., 3)] * mesh.size)
    lats, lons, heights = gridder.regular((-80, 80, -170, 175), (37, 73), z=50e3)    # 2701 points
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    ratio = cases.DEFAULT_RATIO[field]
    try:
        one = forward_raw(field, *cc, pm, ratio)
        for k in (3, 5, 7, 11):
            assert (lons.size // k) % 32 != 0 or k == 3
            many = forward_raw(field, *cc, pm, ratio, devices=[0] * k)
            assert np.array_equal(one, many), (field, k, int(np.sum(one != many)))
        # one NaN (and one huge) coordinate must not change any other point's value
        lon2 = cc[0].copy()
        lon2[40] = np.nan
        lon2[41] = 3.0e6
        bad = forward_raw(field, lon2, cc[1], cc[2], cc[3], pm, ratio)
        keep = np.ones(lons.size, dtype=bool)
        keep[40:42] = False
        assert np.isnan(bad[40])
        assert np.array_equal(bad[keep], one[keep])
    finally:
        pm.close()


@pytest.mark.parametrize("field", ["gz", "potential", "gx", "gzz"])
def test_one_height_specialisation_returns_the_general_bits(field):
    """Points at one height take a specialised sweep (per-piece constants precomputed with the
    same roundings, tvd_far.cu `UR`).  A chunk of a mixed-height point set may be uniform while
    the whole set is not, so the two instantiations must return the same bits -- otherwise
    sharding would change values."""
    from tesseroid_variable_density_b200 import density as D
    from tesseroid_variable_density_b200 import gridder
    from tesseroid_variable_density_b200.mesher import TesseroidMesh
    from tesseroid_variable_density_b200.tesseroid import PreparedModel, _convert_coords, forward_raw
    mesh = TesseroidMesh((-5, 5, -4, 4, 0, -30000.), (1, 16, 20))
    mesh.addprop("density", [D.ExponentialDensity.between(2670, 3300, 0., -30000., 3)] * mesh.size)
    lats, lons, heights = gridder.regular((-8, 8, -9, 9), (20, 30), z=40e3)      # 600 points
    heights = heights.copy()
    heights[300:] = 55e3                                   # second half at another height
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    ratio = cases.DEFAULT_RATIO[field]
    try:
        mixed = forward_raw(field, *cc, pm, ratio)                         # general sweep
        halves = forward_raw(field, *cc, pm, ratio, devices=[0, 0])       # two uniform chunks
        assert np.array_equal(mixed, halves), int(np.sum(mixed != halves))
    finally:
        pm.close()
