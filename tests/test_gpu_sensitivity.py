"""Sensitivity (Jacobian) mode."""
import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import gridder, tesseroid
from tesseroid_variable_density_b200.mesher import TesseroidMesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field", ["gz", "potential", "gy", "gzz"])
def test_rows_sum_to_the_field_and_columns_are_single_tesseroid_runs(field):
    mesh = cases.exponential_case(1e4, 10)[0]
    lats, lons, heights = cases.GRIDS["global2km"]()
    A = tesseroid.sensitivity(field, lons, lats, heights, mesh, delta=0.1)
    assert A.shape == (lons.size, mesh.size)
    full = getattr(tesseroid, field)(lons, lats, heights, mesh, delta=0.1)
    scale = max(np.max(np.abs(full)), np.max(np.abs(A)))
    assert np.max(np.abs(A.sum(axis=1) - full)) <= 1e-12 * scale
    # the reference's way: one call per tesseroid (tesseroid.py:347-349); near- and far-field
    # columns alike are bit-identical
    for t in (0, 17, 40, 71):
        col = getattr(tesseroid, field)(lons, lats, heights, [mesh[t]], delta=0.1)
        assert np.array_equal(A[:, t], col), t


def test_unit_density_sensitivity_is_linear_in_density():
    mesh = TesseroidMesh((-2, 2, -2, 2, 0, -20e3), (2, 4, 4))
    mesh.addprop("density", [2670.0] * mesh.size)
    lats, lons, heights = gridder.regular((-3, 3, -3, 3), (15, 15), z=10e3)
    A = tesseroid.sensitivity("gz", lons, lats, heights, mesh, dens=1.0)
    rho = np.random.default_rng(0).uniform(-500, 500, mesh.size)
    mesh.addprop("density", list(rho))
    want = tesseroid.gz(lons, lats, heights, mesh)
    assert np.max(np.abs(A @ rho - want)) <= 1e-12 * np.max(np.abs(want))


def test_point_slabs_do_not_change_the_rows(monkeypatch):
    """tvd_sensitivity walks large point sets in slabs (device memory for (Q + T) x slab values);
    a row's value at a point must not depend on the slab it fell into."""
    mesh = cases.exponential_case(1e4, 10)[0]
    lats, lons, heights = gridder.regular((-80, 80, 0, 355), (40, 50), z=2000.)     # 2000 points
    whole = tesseroid.sensitivity("gz", lons, lats, heights, mesh, delta=0.1)
    monkeypatch.setenv("TVD_DEBUG_JAC_SLAB", "512")
    slabs = tesseroid.sensitivity("gz", lons, lats, heights, mesh, delta=0.1)
    assert np.array_equal(whole, slabs)
