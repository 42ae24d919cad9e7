"""Host-side logic of the drop-in layer: no GPU needed."""
import doctest
import os
import re
import warnings

import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import (density, gridder, mesher, model, tesseroid)
from tesseroid_variable_density_b200 import _lib
from tesseroid_variable_density_b200.distributed import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_module_constants_match_reference():
    # tesseroid.py:100-103
    assert tesseroid.RATIO_V == 1 and tesseroid.RATIO_G == 2.5
    assert tesseroid.STACK_SIZE == 200 and tesseroid.DELTA == 1e-2


def test_public_signature_matches_reference():
    import inspect
    for name, ratio in (("potential", 1), ("gx", 2.5), ("gy", 2.5), ("gz", 2.5)):
        sig = inspect.signature(getattr(tesseroid, name))
        names = list(sig.parameters)
        assert names[:9] == ["lon", "lat", "height", "model", "dens", "ratio", "njobs", "pool",
                             "delta"]
        assert sig.parameters["ratio"].default == ratio
        assert sig.parameters["delta"].default == 1e-2
        assert sig.parameters["dens"].default is None


def test_split_arrays_doctest():
    res = doctest.run_docstring_examples(tesseroid._split_arrays, {"_split_arrays":
                                                                  tesseroid._split_arrays})
    chunks = tesseroid._split_arrays([[1, 2, 3, 4, 5, 6, 7]], ['meh'], 3)
    assert chunks == [[[1, 2], 'meh'], [[3, 4], 'meh'], [[5, 6, 7], 'meh']]
    assert res is None


@pytest.mark.parametrize("n,world", [(7, 3), (6, 3), (5, 8), (100, 1), (1000001, 8)])
def test_shard_bounds_same_rule_as_split_arrays(n, world):
    if world > n:
        pytest.skip("the reference never makes more chunks than points meaningfully")
    arr = list(range(n))
    chunks = tesseroid._split_arrays([arr], [], world)
    for r in range(world):
        s, e = shard_bounds(n, world, r)
        assert chunks[r][0] == arr[s:e]


def test_check_input_assertions():
    lon = np.zeros(3)
    with pytest.raises(AssertionError):
        tesseroid._check_input(lon, np.zeros(4), lon, None, 1, 1, None)
    with pytest.raises(AssertionError):
        tesseroid._check_input(lon, lon, lon, None, 0, 1, None)
    with pytest.raises(AssertionError):
        tesseroid._check_input(lon, lon, lon, None, 1, 0, None)
    with pytest.raises(AssertionError):
        tesseroid._check_input(lon, lon, lon, None, 1, 1, object())
    assert tesseroid._check_input(lon, lon, lon, None, 1, 1, None).shape == (3,)


def test_convert_coords_is_numpy_identical():
    rng = np.random.default_rng(3)
    lon, lat, h = rng.uniform(0, 360, 50), rng.uniform(-90, 90, 50), rng.uniform(0, 3e5, 50)
    lo, sl, cl, r = tesseroid._convert_coords(lon, lat, h)
    assert np.array_equal(lo, np.radians(lon))
    assert np.array_equal(sl, np.sin(np.radians(lat)))
    assert np.array_equal(cl, np.cos(np.radians(lat)))
    assert np.array_equal(r, 6378137.0 + h)


def test_density_objects_follow_script_expressions():
    h = np.linspace(-1000., 0., 101)
    top, bottom, b = 0., -1000., 30
    A = (3300 - 2670) / (1 - np.exp(-b))
    C = 3300 - A
    d = density.ExponentialDensity.between(2670, 3300, top, bottom, b)
    assert np.array_equal(d(h), A * np.exp(-b * (h - bottom) / (top - bottom)) + C)
    slope, const = -0.63, 4000.
    assert np.array_equal(density.LinearDensity.from_radius(slope, const)(h),
                          slope * (h + 6378137.0) + const)
    k = 2 * np.pi * 5 / 1000.
    assert np.array_equal(density.SineDensity(1650., k, 1650.)(h), 3300. / 2 * np.sin(k * h) + 3300. / 2)
    # scalars too (the reference calls density(rc[k] - R) with a Python float)
    assert d(-500.0) == A * np.exp(-b * (-500.0 - bottom) / (top - bottom)) + C


def test_as_kind_params():
    assert density.as_kind_params(2670) == (density.CONST, (2670.0, 0.0, 0.0, 0.0, 0.0))
    assert density.as_kind_params(np.float64(1.5))[0] == density.CONST
    assert density.as_kind_params(density.SineDensity(1, 2, 3)) == (density.SINE, (1.0, 2.0, 3.0, 0.0, 0.0))
    with pytest.raises(TypeError):
        density.as_kind_params(lambda h: 1.0)        # no CPU fallback for arbitrary callables
    with pytest.raises(TypeError):
        density.as_kind_params("rock")


def test_flatten_generic_iterable_follows_check_tesseroid():
    # tesseroid.py:142-161: None skipped, no density skipped unless dens given, bounds asserted
    t1 = mesher.Tesseroid(0, 1, 0, 1, 0, -100, props={"density": 500})
    t2 = mesher.Tesseroid(1, 2, 0, 1, 0, -100)
    t3 = mesher.Tesseroid(2, 3, 0, 1, 0, -100, props={"density": density.LinearDensity(1, 2, 3)})
    flat = model.flatten_model([t1, None, t2, t3])
    assert flat.size == 2
    assert flat.kind.tolist() == [density.CONST, density.LINEAR]
    assert flat.bounds[1].tolist() == [2, 3, 0, 1, 0, -100]
    flat = model.flatten_model([t1, None, t2, t3], dens=1000.)
    assert flat.size == 3 and np.all(flat.kind == density.CONST) and np.all(flat.params[:, 0] == 1000.)
    bad = mesher.Tesseroid(1, 0, 0, 1, 0, -100, props={"density": 1})
    with pytest.raises(AssertionError, match="Invalid tesseroid dimensions"):
        model.flatten_model([bad])
    bad = mesher.Tesseroid(0, 1, 0, 1, -100, 0, props={"density": 1})
    with pytest.raises(AssertionError):
        model.flatten_model([bad])


def test_flatten_generator_with_short_lived_density_objects():
    # a generator that builds a fresh density per tesseroid: object ids are recycled as soon as
    # the previous tesseroid dies, so the per-object cache must not key on a dead object's id
    def cells():
        for i in range(200):
            yield mesher.Tesseroid(i, i + 1, 0, 1, 0, -100,
                                   props={"density": density.LinearDensity(float(i), 2., 3.)})
    flat = model.flatten_model(cells())
    assert flat.size == 200
    assert flat.params[:, 0].tolist() == [float(i) for i in range(200)]


def test_flatten_mesh_fast_path_equals_iteration():
    mesh = mesher.TesseroidMesh((-1, 1, -1, 1, 0, -35000), (2, 3, 3))
    dens = density.ExponentialDensity(1., -1., 0., 1e5, 2.)
    lst = [dens] * mesh.size
    lst[0] = 500
    lst[1] = density.LinearDensity(-0.01, 0., 0.)
    mesh.addprop("density", lst)
    fast = model.flatten_model(mesh)
    slow = model.flatten_model([c for c in mesh])
    assert np.array_equal(fast.bounds, slow.bounds)
    assert np.array_equal(fast.kind, slow.kind)
    assert np.array_equal(fast.params, slow.params)
    mesh.mask = [4]
    assert model.flatten_model(mesh).size == mesh.size - 1
    # numeric property arrays are constant densities
    mesh2 = mesher.TesseroidMesh((-1, 1, -1, 1, 0, -35000), (1, 2, 2))
    mesh2.addprop("density", np.array([1., 2., 3., 4.]))
    f2 = model.flatten_model(mesh2)
    assert np.all(f2.kind == density.CONST) and f2.params[:, 0].tolist() == [1., 2., 3., 4.]
    # a mesh without densities contributes nothing unless dens is given
    mesh3 = mesher.TesseroidMesh((-1, 1, -1, 1, 0, -35000), (1, 2, 2))
    assert model.flatten_model(mesh3).size == 0
    assert model.flatten_model(mesh3, dens=2.0).size == 4


def test_density_table_is_columnar_sequence():
    params = np.zeros((3, 5))
    params[:, 0] = [1, 2, 3]
    params[:, 1] = 0.5
    tab = model.DensityTable(density.SINE, params)
    assert len(tab) == 3 and isinstance(tab[1], density.SineDensity) and tab[1].amplitude == 2
    mesh = mesher.TesseroidMesh((0, 3, 0, 1, 0, -10), (1, 1, 3))
    mesh.addprop("density", tab)
    flat = model.flatten_model(mesh)
    assert np.all(flat.kind == density.SINE) and np.array_equal(flat.params, params)


def test_tesseroid_model_matches_reference_arithmetic():
    # tesseroid_model.py:139-154 including the `s = lats[i] - dlon/2.` quirk
    nlat, nlon = 4, 5
    top = np.arange(20, dtype=float)
    bottom = top - 100.
    tm = mesher.TesseroidModel((-40., -39., 288., 290.), top, bottom, (nlat, nlon))
    dlat, dlon = tm.spacing
    cell = tm[7]
    i, j = 1, 2
    assert cell.w == tm.lons[j] - dlon / 2. and cell.e == cell.w + dlon
    assert cell.s == tm.lats[i] - dlon / 2. and cell.n == cell.s + dlat
    flat, _ = tm.flat_bounds()
    assert np.array_equal(flat, np.array([c.get_bounds() for c in tm]))
    # inverted cells are flipped (tesseroid_model.py:102-110)
    tm2 = mesher.TesseroidModel((-40., -39., 288., 290.), bottom, top, (nlat, nlon))
    assert np.array_equal(tm2.top, top) and np.array_equal(tm2.bottom, bottom)


def test_mesh_and_grid_match_fatiando_shims():
    import sys
    shims = os.path.join(ROOT, "oracle", "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    from fatiando import mesher as fm, gridder as fg       # oracle/shims
    a = mesher.TesseroidMesh((0, 360, -90, 90, 0, -1000.), (1, 6, 12))
    b = fm.TesseroidMesh((0, 360, -90, 90, 0, -1000.), (1, 6, 12))
    assert [c.get_bounds() for c in a] == [c.get_bounds() for c in b]
    for x, y in zip(gridder.regular((-90, 90, 0, 360), (19, 13), z=2e3),
                    fg.regular((-90, 90, 0, 360), (19, 13), z=2e3)):
        assert np.array_equal(x, y)


def test_product_never_touches_the_oracle():
    """No file of the product package may import, load or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "tesseroid_variable_density_b200")
    pat = re.compile(r"oracle|liboracle|build_ref|tess_oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), "%s mentions the oracle" % f


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tvd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tvd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), "libtvd.so does not export %s" % n
        assert n in _lib.SIGNATURES, "%s has no ctypes signature" % n
    assert sorted(_lib.SIGNATURES) == names


def test_fails_loudly_without_a_gpu():
    lib = _lib.load()
    if lib.tvd_device_count() > 0:
        pytest.skip("a GPU is present")
    mesh = mesher.TesseroidMesh((0, 1, 0, 1, 0, -100), (1, 1, 1))
    mesh.addprop("density", [1.0])
    with pytest.raises(RuntimeError, match="no CPU fallback|no such CUDA device"):
        tesseroid.gz(np.zeros(1), np.zeros(1), np.full(1, 1e3), mesh)


def test_analytical_shells_consistent():
    # gz = V/r, gzz = 2V/r^2, gxx = gyy = -V/r^2 (linear_density.py:20-30)
    from tesseroid_variable_density_b200 import analytical
    d = analytical.shell_exponential_density(2000., 0., -1e4, 700., 5, 2600.)
    r = 6378137.0 + 2000.
    assert np.isclose(d["gz"], 1e5 * d["potential"] / r, rtol=1e-15)
    assert np.isclose(d["gzz"], -2 * d["gxx"], rtol=1e-15) and d["gxx"] == d["gyy"]
    # closed forms against a brute-force radial quadrature of V = 4 pi G / r * int rho r'^2 dr'
    from tesseroid_variable_density_b200.constants import G
    top, bottom, height = 0., -1e5, 2000.
    rr = np.linspace(bottom, top, 200001) + 6378137.0
    for dens, exact in (
            (density.ExponentialDensity(700., -5., bottom, top - bottom, 2600.),
             analytical.shell_exponential_density(height, top, bottom, 700., 5, 2600.)),
            (density.SineDensity(1650., 2*np.pi*3/1e5, 1650.),
             analytical.shell_sine_density(height, top, bottom, 1650., 2*np.pi*3/1e5, 1650.)),
            (density.LinearDensity.from_radius(-0.0063, 42000.),
             analytical.shell_linear_density(height, top, bottom, -0.0063, 42000.))):
        f = dens(rr - 6378137.0) * rr**2
        integral = np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(rr))
        v = 4 * np.pi * G * integral / (6378137.0 + height)
        assert np.isclose(v, exact["potential"], rtol=1e-8)


def test_sediment_thickness_reader_and_neuquen_layer():
    """datasets.load_sediment_thickness reads data/sediment_thickness.dat the way the reference
    does (numpy.loadtxt(..., unpack=True), data/README.md:25-37, neuquen_basin.py:33-47) and
    neuquen_basin_model builds the layer of neuquen_basin.py:55-72; both must agree with the
    fixture the golden replays use."""
    import cases
    from tesseroid_variable_density_b200 import datasets
    sed = datasets.load_sediment_thickness()
    fix = np.load(os.path.join(cases.GOLDEN, "neuquen_thickness.npz"))
    assert sed["shape"] == (117, 91) and sed["lat"].size == 117 * 91
    for k in ("lat", "lon", "thickness"):
        assert np.array_equal(sed[k], fix[k], equal_nan=True)
    assert sed["area"] == (sed["lat"].min(), sed["lat"].max(), sed["lon"].min(), sed["lon"].max())
    basin, top, bottom = datasets.neuquen_basin_model()
    basin2, top2, bottom2 = cases.neuquen_model()
    assert (top, bottom) == (top2, bottom2)
    assert np.array_equal(basin.flat_bounds()[0], basin2.flat_bounds()[0])
    nans = np.isnan(sed["thickness"])
    assert int(nans.sum()) == 5692          # degenerate top = bottom = 0 cells, still evaluated
    assert np.all(basin.top[nans] == 0) and np.all(basin.bottom[nans] == 0)
