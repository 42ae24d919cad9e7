"""
Replay of the reference's stored golden results (tests/golden/reference_results) through the
CUDA path and its public Python surface.  The stored numbers are 100*max|rel. difference to the
analytical shell|; the CUDA field agrees with the reference's to ~1e-10 relative at 260 km and
to the conditioning limit at z = 0, which bounds how far the stored percentages can move.
"""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

# (family, file, sweep points, absolute tolerance on the stored percentage)
GOLDENS = [
    ("linear", "potential-260km-1000.npz", 10, 1e-8),
    ("linear", "gz-260km-100000.npz", 10, 1e-8),
    ("linear", "gz-260km-1000000.npz", 10, 1e-8),
    ("linear", "potential-global-1000.npz", 10, 1e-6),
    ("linear", "gz-global-10000.npz", 10, 1e-4),
    ("linear", "gz-pole-100000.npz", 10, 1e-4),
    ("linear", "potential-equator-1000000.npz", 10, 1e-6),
    ("exponential", "gz-260km-1000000-100.npz", 9, 1e-8),
    ("exponential", "potential-260km-1000-30.npz", 9, 1e-8),
    ("exponential", "gz-260km-10000-5.npz", 9, 1e-8),
    ("exponential", "potential-global-1000-30.npz", 9, 1e-6),
    ("exponential", "gz-global-100000-10.npz", 9, 1e-4),
    ("sine", "potential-260km-1000-10.npz", 5, 1e-8),
    ("sine", "gz-260km-100000-5.npz", 5, 1e-8),
    ("sine", "gz-260km-1000000-1.npz", 5, 1e-8),
    ("grid-search", "potential-global-10000-30.npz", 40, 1e-6),
    ("grid-search", "gz-global-1000000-30.npz", 40, 1e-4),
]


@pytest.mark.parametrize("family,fname,nsweep,atol", GOLDENS)
def test_cuda_reproduces_stored_differences(family, fname, nsweep, atol):
    got, want = cases.replay_golden(cases.CudaEngine(), family, fname, max_sweep=nsweep)
    assert np.max(np.abs(got - want)) <= atol, (got, want)


@pytest.mark.parametrize("kind,field", [("homogeneous", "gz"), ("homogeneous", "potential"),
                                        ("linear", "gz"), ("linear", "potential"),
                                        ("exponential", "gz"), ("exponential", "potential")])
def test_cuda_reproduces_neuquen_grids(kind, field):
    """BASELINE config 3, the full 6 399-point grids of the Neuquen application: <= 1e-9 of the
    reference's stored results."""
    from tesseroid_variable_density_b200 import tesseroid
    stored = np.load(os.path.join(cases.REF_RESULTS, "neuquen", "%s-%s.npz" % (kind, field)))
    basin, top, bottom = cases.neuquen_model()
    basin.addprop("density", [cases.neuquen_densities(top, bottom)[kind]] * basin.size)
    got = getattr(tesseroid, field)(stored["lon"], stored["lat"], stored["height"], basin)
    want = stored["result"]
    assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want))
