"""
The reference's import paths (code/scripts/linear_density.py:5-11) resolve through compat/ and
its scripts' own density callables are recognised (density.fit_density).  The CPU part checks
imports, fitting and model adaptation; the GPU part runs the body of a reference script.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def compat_path():
    p = os.path.join(ROOT, "compat")
    saved = list(sys.path)
    # the reference build of the oracle uses its own fatiando shims: keep the two apart
    for name in [m for m in sys.modules if m == "fatiando" or m.startswith("fatiando.")]:
        sys.modules.pop(name)
    sys.path.insert(0, p)
    yield p
    sys.path[:] = saved
    for name in [m for m in sys.modules if m == "fatiando" or m.startswith("fatiando.")
                 or m == "tesseroid_density" or m.startswith("tesseroid_density.")]:
        sys.modules.pop(name)


def _script_header():
    from fatiando.constants import G, MEAN_EARTH_RADIUS, SI2MGAL, SI2EOTVOS
    from fatiando.mesher import TesseroidMesh
    from fatiando import gridder
    from tesseroid_density import tesseroid
    return G, MEAN_EARTH_RADIUS, SI2MGAL, SI2EOTVOS, TesseroidMesh, gridder, tesseroid


def test_reference_import_block_resolves(compat_path):
    G, R, SI2MGAL, SI2EOTVOS, TesseroidMesh, gridder, tesseroid = _script_header()
    assert (G, R, SI2MGAL, SI2EOTVOS) == (6.673e-11, 6378137.0, 1e5, 1e9)
    for name in ("potential", "gx", "gy", "gz", "RATIO_V", "RATIO_G", "STACK_SIZE", "DELTA",
                 "_density_based_discretization", "_check_input", "_convert_coords",
                 "_split_arrays"):
        assert hasattr(tesseroid, name), name
    assert (tesseroid.RATIO_V, tesseroid.RATIO_G, tesseroid.STACK_SIZE, tesseroid.DELTA) == \
        (1, 2.5, 200, 1e-2)


def test_script_callables_are_recognised(compat_path):
    from tesseroid_variable_density_b200 import density as D
    from tesseroid_density.tesseroid import _adapt
    G, R, SI2MGAL, SI2EOTVOS, TesseroidMesh, gridder, tesseroid = _script_header()
    top, bottom = 0., -1000.
    model = TesseroidMesh((0, 360, -90, 90, top, bottom), (1, 6, 12))
    slope = (2670 - 3300) / (top - bottom)
    constant_term = 2670 - slope * (top + R)

    def density_linear(height):             # linear_density.py:73-75
        r = height + R
        return slope * r + constant_term
    model.addprop("density", [density_linear for i in range(model.size)])
    adapted, _ = _adapt(model, None)
    fitted = adapted.props["density"]
    assert len(fitted) == 72 and all(f is fitted[0] for f in fitted)
    assert isinstance(fitted[0], D.LinearDensity)
    h = np.linspace(bottom, top, 33)
    assert np.max(np.abs(fitted[0](h) - density_linear(h))) <= 1e-9 * 3300

    b_factor = 30.
    A = (3300 - 2670) / (1 - np.exp(-b_factor))

    def density_exponential(height):        # exponential_density.py:97-98
        return A * np.exp(-b_factor * (height - bottom) / (top - bottom)) + (3300 - A)
    f = D.fit_density(density_exponential, top, bottom)
    assert isinstance(f, D.ExponentialDensity)
    assert np.max(np.abs(f(h) - density_exponential(h))) <= 1e-9 * 3300
    f = D.fit_density(lambda hh: 1650 * np.sin(2 * np.pi * 5 * hh / 1000.) + 1650, top, bottom)
    assert isinstance(f, D.SineDensity)
    assert D.fit_density(lambda hh: 2670. + 0 * hh, top, bottom) == 2670.
    with pytest.raises(TypeError):
        D.fit_density(lambda hh: hh ** 2, top, bottom)


@pytest.mark.gpu
def test_reference_script_body_runs_unchanged(compat_path):
    """The computation of code/scripts/linear_density.py:44-100 (one thickness, one grid, one D)
    written exactly as the script writes it -- Python callable included."""
    from tesseroid_variable_density_b200 import analytical
    G, R, SI2MGAL, SI2EOTVOS, TesseroidMesh, gridder, tesseroid = _script_header()
    thickness = 1000.
    top, bottom = 0, -thickness
    model = TesseroidMesh((0, 360, -90, 90, top, bottom), (1, 6, 12))
    lats, lons, heights = gridder.regular((-90, 90, 0, 360), (19, 13), z=2000.)
    slope = (2670 - 3300) / (top - bottom)
    constant_term = 2670 - slope * (top + R)

    def density_linear(height):
        r = height + R
        return slope * r + constant_term
    model.addprop("density", [density_linear for i in range(model.size)])
    analytical_value = analytical.shell_linear_density(heights[0], top, bottom, slope,
                                                       constant_term)
    for field, D_ratio in (("potential", 1), ("gz", 2.5)):
        result = getattr(tesseroid, field)(lons, lats, heights, model, ratio=D_ratio, delta=None)
        diff = np.abs((result - analytical_value[field]) / analytical_value[field])
        assert 100 * np.max(diff) <= 0.1        # the paper's 0.1 %
