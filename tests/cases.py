"""
Shared definitions of the reference's validation cases (de-facto tests of the reference:
code/scripts/linear_density.py, exponential_density.py, sine_density.py, grid_search.py,
neuquen_basin.py) and small helpers to run them through either the oracle or the CUDA path.
"""
from __future__ import annotations

import os

import numpy as np

from tesseroid_variable_density_b200 import analytical, gridder
from tesseroid_variable_density_b200.constants import G, MEAN_EARTH_RADIUS, SI2EOTVOS, SI2MGAL
from tesseroid_variable_density_b200.density import (ExponentialDensity, LinearDensity,
                                                     SineDensity)
from tesseroid_variable_density_b200.mesher import TesseroidMesh, TesseroidModel
from tesseroid_variable_density_b200.model import flatten_model
from tesseroid_variable_density_b200.tesseroid import _convert_coords

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
REF_RESULTS = os.path.join(GOLDEN, "reference_results")

ALL_FIELDS = ("potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz")
DEFAULT_RATIO = {"potential": 1, "gx": 2.5, "gy": 2.5, "gz": 2.5, "gxx": 8, "gxy": 8, "gxz": 8,
                 "gyy": 8, "gyz": 8, "gzz": 8}
UNIT = {"potential": G, "gx": G*SI2MGAL, "gy": G*SI2MGAL, "gz": G*SI2MGAL}
for _f in ALL_FIELDS[4:]:
    UNIT[_f] = G*SI2EOTVOS

# the four grids of the shell scripts (linear_density.py:54-58) plus the 2 km variants that
# BASELINE.json config 1 uses
GRIDS = {
    "pole": lambda: gridder.regular((89, 90, 0, 1), (11, 11), z=0),
    "equator": lambda: gridder.regular((0, 1, 0, 1), (11, 11), z=0),
    "global": lambda: gridder.regular((-90, 90, 0, 360), (19, 13), z=0),
    "260km": lambda: gridder.regular((-90, 90, 0, 360), (19, 13), z=260e3),
    "global2km": lambda: gridder.regular((-90, 90, 0, 360), (19, 13), z=2000.),
    "equator2km": lambda: gridder.regular((0, 1, 0, 1), (11, 11), z=2000.),
    "pole2km": lambda: gridder.regular((89, 90, 0, 1), (11, 11), z=2000.),
}


def shell_mesh(thickness):
    """72-tesseroid spherical shell (linear_density.py:44-49)."""
    return TesseroidMesh((0, 360, -90, 90, 0, -thickness), (1, 6, 12))


def linear_case(thickness):
    """density + analytical function of linear_density.py:66-75."""
    mesh = shell_mesh(thickness)
    top, bottom = mesh.bounds[4], mesh.bounds[5]
    slope = (2670 - 3300) / (top - bottom)
    constant_term = 2670 - slope * (top + MEAN_EARTH_RADIUS)
    dens = LinearDensity.from_radius(slope, constant_term)
    mesh.addprop("density", [dens] * mesh.size)

    def exact(height):
        return analytical.shell_linear_density(height, top, bottom, slope, constant_term)
    return mesh, dens, exact


def exponential_case(thickness, b_factor):
    """exponential_density.py:87-103."""
    mesh = shell_mesh(thickness)
    top, bottom = mesh.bounds[4], mesh.bounds[5]
    density_bottom, density_top = 3300, 2670
    A = (density_bottom - density_top) / (1 - np.exp(-b_factor))
    C = density_bottom - A
    dens = ExponentialDensity(A, -b_factor, bottom, top - bottom, C)
    mesh.addprop("density", [dens] * mesh.size)

    def exact(height):
        return analytical.shell_exponential_density(height, top, bottom, A, b_factor, C)
    return mesh, dens, exact


def sine_case(thickness, b_factor):
    """sine_density.py:82-100."""
    mesh = shell_mesh(thickness)
    top, bottom = mesh.bounds[4], mesh.bounds[5]
    max_density = 3300.
    k_factor = 2 * np.pi * b_factor / (top - bottom)
    dens = SineDensity(max_density / 2, k_factor, max_density / 2)
    mesh.addprop("density", [dens] * mesh.size)

    def exact(height):
        return analytical.shell_sine_density(height, top, bottom, max_density/2, k_factor,
                                             max_density/2)
    return mesh, dens, exact


def percent_difference(result, exact_value):
    """The scripts' figure of merit (linear_density.py:95-96)."""
    diff = np.abs((result - exact_value) / exact_value)
    return 100 * np.max(diff)


class OracleEngine(object):
    """field(lons, lats, heights, mesh, ratio, delta) in output units through the C oracle."""

    def __init__(self, oracle, n_threads=None):
        self.oracle = oracle
        self.n_threads = n_threads or min(8, os.cpu_count() or 1)

    def __call__(self, field, lons, lats, heights, mesh, ratio, delta):
        flat = flatten_model(mesh)
        cc = _convert_coords(np.asarray(lons, float), np.asarray(lats, float),
                             np.asarray(heights, float))
        res, _ = self.oracle.forward(field, flat.bounds, flat.kind, flat.params, delta, ratio,
                                     *cc, n_threads=self.n_threads)
        res *= UNIT[field]
        return res


class CudaEngine(object):
    """Same call through the drop-in Python surface of the CUDA library."""

    def __call__(self, field, lons, lats, heights, mesh, ratio, delta):
        from tesseroid_variable_density_b200 import tesseroid
        return getattr(tesseroid, field)(lons, lats, heights, mesh, ratio=ratio, delta=delta)


class ReferenceEngine(object):
    """Same call through the unmodified reference (oracle/_ref)."""

    def __init__(self, ref):
        self.ref = ref

    def __call__(self, field, lons, lats, heights, mesh, ratio, delta):
        return getattr(self.ref, field)(lons, lats, heights, mesh, ratio=ratio, delta=delta)


def replay_golden(engine, family, fname, max_sweep=None):
    """
    Recompute one stored result file of the reference (tests/golden/reference_results/<family>)
    with `engine`.  Returns (computed differences, stored differences) as flat arrays
    (possibly truncated to the first `max_sweep` sweep points).
    """
    stored = np.load(os.path.join(REF_RESULTS, family, fname))
    parts = fname[:-4].split("-")
    field, grid_name, thickness = parts[0], parts[1], float(parts[2])
    lats, lons, heights = GRIDS[grid_name]()
    out = []
    if family == "linear":
        mesh, _, exact = linear_case(thickness)
        ex = exact(heights[0])[field]
        sweep = [(D, None) for D in stored["D_values"]]
    elif family == "exponential":
        mesh, _, exact = exponential_case(thickness, int(parts[3]))
        ex = exact(heights[0])[field]
        sweep = [(DEFAULT_RATIO[field], d) for d in stored["delta_values"]]
    elif family == "sine":
        mesh, _, exact = sine_case(thickness, int(parts[3]))
        ex = exact(heights[0])[field]
        sweep = [(DEFAULT_RATIO[field], d) for d in stored["delta_values"]]
    elif family == "grid-search":
        mesh, _, exact = exponential_case(thickness, int(parts[3]))
        ex = exact(heights[0])[field]
        sweep = [(D, d) for d, D in zip(stored["delta_grid"].ravel(), stored["D_grid"].ravel())]
    else:
        raise ValueError(family)
    want = stored["differences"].ravel()
    if max_sweep is not None:
        sweep, want = sweep[:max_sweep], want[:max_sweep]
    for ratio, delta in sweep:
        res = engine(field, lons, lats, heights, mesh, ratio, delta)
        out.append(percent_difference(res, ex))
    return np.array(out), want


# ---- Neuquen basin application (neuquen_basin.py) -----------------------------------------
#: constant top of the sediment layer; data/topography.npy is missing from the reference
#: (.MISSING_LARGE_BLOBS) and this value reproduces its stored results
NEUQUEN_TOP = 845.0761706691


def neuquen_model(thickness_file=None):
    """TesseroidModel of neuquen_basin.py:33-72 from the committed thickness fixture."""
    path = thickness_file or os.path.join(GOLDEN, "neuquen_thickness.npz")
    d = np.load(path)
    lat, lon, thickness = d["lat"], d["lon"], d["thickness"]
    shape = (117, 91)
    area = (lat.min(), lat.max(), lon.min(), lon.max())
    nans = np.isnan(thickness)
    top = np.full(thickness.shape, NEUQUEN_TOP)
    top[nans] = np.nan
    bottom = top - thickness
    t, b = top.copy(), bottom.copy()
    t[nans] = 0
    b[nans] = 0
    basin = TesseroidModel(area, t, b, shape)
    stop, sbottom = basin.top[~nans].max(), basin.bottom[~nans].min()
    return basin, stop, sbottom


def neuquen_densities(top, bottom):
    """homogeneous / linear / exponential densities of neuquen_basin.py:90-147."""
    density_top, density_bottom = -412, -275
    mean_density = (density_top + density_bottom) / 2
    linear = LinearDensity.between(density_top, density_bottom, top, bottom)
    expo = ExponentialDensity.between(density_top, density_bottom, top, bottom, 3)
    return {"homogeneous": mean_density, "linear": linear, "exponential": expo}
