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
