"""The drop-in Python surface on the GPU: same behaviour as tesseroid_density.tesseroid."""
import warnings

import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import density as D
from tesseroid_variable_density_b200 import gridder, tesseroid, _lib
from tesseroid_variable_density_b200.mesher import Tesseroid, TesseroidMesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field", cases.ALL_FIELDS)
def test_analytical_shell_all_ten_fields(field):
    """<= 0.1 % of the analytical shell at the paper's defaults (manuscript.tex:816-820);
    components that vanish by symmetry are compared with the scale of gz / gzz."""
    mesh, _, exact = cases.exponential_case(1e4, 10)
    lats, lons, heights = cases.GRIDS["260km"]()
    got = getattr(tesseroid, field)(lons, lats, heights, mesh, delta=0.1)
    ex = exact(heights[0])
    want = ex[field]
    if want != 0:
        assert cases.percent_difference(got, want) <= 0.1
    else:
        scale = abs(ex["gz"]) if field in ("gx", "gy") else abs(ex["gzz"])
        assert np.max(np.abs(got)) <= 1e-3 * scale


@pytest.mark.parametrize("case", ["linear", "exponential", "sine"])
@pytest.mark.parametrize("grid", ["global2km", "pole2km", "equator2km"])
def test_analytical_shell_at_2km(case, grid):
    # BASELINE config 1: gz + potential at 2 km height vs the analytical solution
    if case == "linear":
        mesh, _, exact = cases.linear_case(1e4)
        delta = None
    elif case == "exponential":
        mesh, _, exact = cases.exponential_case(1e4, 5)
        delta = 0.1
    else:
        mesh, _, exact = cases.sine_case(1e4, 2)
        delta = 0.1
    lats, lons, heights = cases.GRIDS[grid]()
    ex = exact(heights[0])
    for field in ("potential", "gz"):
        got = getattr(tesseroid, field)(lons, lats, heights, mesh, delta=delta)
        assert cases.percent_difference(got, ex[field]) <= 0.1


def test_density_based_discretization_contract():
    # same return contract as tesseroid.py:237-283: list of [w, e, s, n, top, bottom] arrays
    A = 1 / (1 - np.exp(-5))
    dens = D.ExponentialDensity(A, -5, -1000., 1000., 1 - A)
    subset = tesseroid._density_based_discretization([-10, 10, -10, 10, 0, -1000], dens, 0.1)
    assert [list(np.round(s, 6)) for s in subset] == [[-10, 10, -10, 10, 0, -680],
                                                      [-10, 10, -10, 10, -680, -1000]]
    sine = D.SineDensity(0.5, 2 * np.pi * 2 / 1000., 0.5)
    subset = tesseroid._density_based_discretization([-10, 10, -10, 10, 0, -1000], sine, 0.1)
    assert len(subset) == 5
    # FIFO emission order, not depth order
    assert round(subset[0][4], 1) == -384.3 and round(subset[3][4], 1) == 0.0
    lin = D.LinearDensity(-0.01, 0., 0.)
    assert len(tesseroid._density_based_discretization([-10, 10, -10, 10, 0, -1000], lin, 1e-4)) == 1


def test_k1_matches_oracle_on_many_tesseroids(oracle_lib):
    rng = np.random.default_rng(5)
    n = 500
    tops = rng.uniform(-1000, 2000, n)
    bots = tops - rng.uniform(0, 60000, n)
    bots[::50] = tops[::50]                   # degenerate cells
    bounds = np.zeros((n, 6))
    bounds[:, 1] = 1
    bounds[:, 3] = 1
    bounds[:, 4], bounds[:, 5] = tops, bots
    kinds = rng.integers(0, 4, n).astype(np.int32)
    params = np.zeros((n, 5))
    for i in range(n):
        if kinds[i] == D.CONST:
            params[i, 0] = 2670
        elif kinds[i] == D.LINEAR:
            params[i, :3] = (-0.01, 6378137.0, 60000.)
        elif kinds[i] == D.EXP:
            params[i] = (600., -rng.integers(1, 60), bots[i], max(tops[i] - bots[i], 1.0), 2600.)
        else:
            params[i, :3] = (1650., 2 * np.pi * rng.integers(1, 6) / max(tops[i] - bots[i], 1.0), 1650.)
    from tesseroid_variable_density_b200.model import FlatModel
    for delta in (0.1, 1e-2, 1e-3):
        pm = tesseroid.PreparedModel(FlatModel(bounds, kinds, params), delta=delta)
        t, b, parent = pm.pieces()
        pm.close()
        pos = 0
        for i in range(n):
            if kinds[i] == D.CONST:
                ot, ob = np.array([tops[i]]), np.array([bots[i]])
            else:
                ot, ob = oracle_lib.density_discretize(tops[i], bots[i], kinds[i], params[i], delta)
            k = len(ot)
            assert np.array_equal(parent[pos:pos + k], np.full(k, i))
            assert np.array_equal(t[pos:pos + k], ot) and np.array_equal(b[pos:pos + k], ob), (i, delta)
            pos += k
        assert pos == t.size


def test_njobs_dens_delta_none_and_shapes(oracle_lib):
    mesh, dens, _ = cases.exponential_case(1e4, 5)
    lats, lons, heights = cases.GRIDS["global2km"]()
    base = tesseroid.gz(lons, lats, heights, mesh, delta=0.1)
    # njobs maps to the number of GPUs; values do not depend on it (tesseroid.py:179-195)
    again = tesseroid.gz(lons, lats, heights, mesh, delta=0.1, njobs=4)
    assert np.array_equal(base, again)
    # 2-D input arrays keep their shape (zeros_like(lon), tesseroid.py:121)
    shaped = tesseroid.gz(lons.reshape(19, 13), lats.reshape(19, 13), heights.reshape(19, 13), mesh, delta=0.1)
    assert shaped.shape == (19, 13) and np.array_equal(shaped.ravel(), base)
    # dens overrides the model's density (tesseroid.py:157-158)
    c1 = tesseroid.gz(lons, lats, heights, mesh, dens=1000.)
    c2 = tesseroid.gz(lons, lats, heights, mesh, dens=2000.)
    assert np.allclose(c2, 2 * c1, rtol=1e-13)
    # delta=None disables the radial splitting (tesseroid.py:222)
    pm = tesseroid.PreparedModel(mesh, delta=None)
    assert pm.n_pieces == mesh.size
    pm.close()
    pm = tesseroid.PreparedModel(mesh, delta=0.1)
    assert pm.n_pieces == 2 * mesh.size
    one = tesseroid.gz(lons, lats, heights, pm)         # a prepared model is accepted as `model`
    assert np.array_equal(one, base)
    pm.close()
    with pytest.raises(AssertionError):
        tesseroid.gz(lons, lats[:-1], heights, mesh)
    with pytest.raises(AssertionError):
        tesseroid.gz(lons, lats, heights, mesh, ratio=0)
    with pytest.raises(TypeError):
        tesseroid.gz(lons, lats, heights, mesh, dens=lambda h: 1.0)


def test_empty_inputs():
    mesh, _, _ = cases.exponential_case(1e4, 5)
    empty = np.zeros(0)
    assert tesseroid.gz(empty, empty, empty, mesh).shape == (0,)
    lats, lons, heights = cases.GRIDS["pole2km"]()
    assert np.all(tesseroid.gz(lons, lats, heights, []) == 0)
    assert np.all(tesseroid.potential(lons, lats, heights, [None, Tesseroid(0, 1, 0, 1, 0, -1)]) == 0)


def test_too_small_warning_and_stack_overflow(monkeypatch):
    # a point inside a microscopic cell: the reference stops dividing below 0.1 m and warns
    # (tesseroid.py:212-216; _tesseroid.pyx:136-137)
    cell = [Tesseroid(0, 1e-7, 0, 1e-7, 0, -1e-3, props={"density": 1000.})]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        tesseroid.gz(np.array([5e-8]), np.array([5e-8]), np.array([1e-4]), cell)
    assert any(issubclass(x.category, RuntimeWarning) and "Stopped dividing" in str(x.message) for x in w)
    # stack overflow -> ValueError (_tesseroid.pyx:88-89)
    monkeypatch.setattr(tesseroid, "STACK_SIZE", 3)
    big = [Tesseroid(-10, 10, -10, 10, 0, -1000, props={"density": 1000.})]
    with pytest.raises(ValueError, match="Tesseroid stack overflow"):
        tesseroid.gz(np.array([0.]), np.array([0.]), np.array([10.]), big)


def test_own_trig_against_glibc():
    """The sweep's cos/sin vs math.cos/math.sin (glibc, what the reference calls): <= 1 ulp and
    equal in >= 95 % of cases; >= 98 % for the small arguments of regional models."""
    import math
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for lim, frac in ((0.35, 0.98), (0.785, 0.95), (7.0, 0.95), (1000.0, 0.95)):
        x = rng.uniform(-lim, lim, 100000)
        c, s = np.empty_like(x), np.empty_like(x)
        assert lib.tvd_debug_trig(0, x.size, _lib.ptr_f64(x), _lib.ptr_f64(c), _lib.ptr_f64(s)) == 0
        cref = np.array([math.cos(v) for v in x])
        sref = np.array([math.sin(v) for v in x])
        assert np.max(np.abs(c - cref) / np.spacing(np.abs(cref))) <= 1.0
        assert np.max(np.abs(s - sref) / np.spacing(np.abs(sref))) <= 1.0
        assert np.mean(c == cref) >= frac and np.mean(s == sref) >= frac


def test_double_double_math_is_correctly_rounded():
    """K1's sin/cos/exp (tvd_ddmath.cuh) against mpmath at 60 digits: correctly rounded for every
    sample; hence equal to glibc wherever glibc itself rounds correctly (>= 99 %)."""
    import math
    import mpmath
    mpmath.mp.dps = 60
    lib = _lib.load()
    rng = np.random.default_rng(11)
    sets = {
        0: (np.concatenate([rng.uniform(-8, 8, 3000), rng.uniform(-2000, 2000, 2000),
                            np.array([0.0, 1e-300, 1e-9, np.pi, 0.5 * np.pi, 710.0, 1e5])]),
            mpmath.sin, math.sin),
        1: (np.concatenate([rng.uniform(-8, 8, 3000), rng.uniform(-2000, 2000, 2000),
                            np.array([0.0, 1e-9, np.pi, 0.5 * np.pi, 1e5])]), mpmath.cos, math.cos),
        2: (np.concatenate([rng.uniform(-40, 5, 4000), rng.uniform(-690, 690, 1000),
                            np.array([0.0, 1e-17, -1e-17, 1.0, -1.0])]), mpmath.exp, math.exp),
    }
    for func, (x, exact, libm) in sets.items():
        x = np.ascontiguousarray(x)
        out = np.empty_like(x)
        assert lib.tvd_debug_math(0, func, x.size, _lib.ptr_f64(x), _lib.ptr_f64(out)) == 0
        want = np.array([float(exact(mpmath.mpf(float(v)))) for v in x])   # float() rounds to nearest
        assert np.array_equal(out, want), (func, int(np.sum(out != want)))
        glibc = np.array([libm(float(v)) for v in x])
        assert np.mean(out == glibc) >= 0.99


def test_table_trig_equals_double_double():
    """The near-field walk's sin/cos (tvd_crtrig.cuh: table + Ziv's rounding test, double-double
    only when the test fails) must return the bits of the double-double versions, i.e. the
    correctly rounded values -- what glibc handed the reference."""
    import mpmath
    mpmath.mp.dps = 60
    lib = _lib.load()
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-0.8, 0.8, 400000), rng.uniform(-1.6, 1.6, 400000),
                        rng.uniform(-7, 7, 400000), rng.uniform(-400, 400, 200000),
                        rng.uniform(-1e-3, 1e-3, 100000), rng.uniform(-1e-9, 1e-9, 1000),
                        np.array([0.0, -0.0, 1e-300, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi,
                                  0.78539816339744828, 0.78539816339744839, 1e5, 9.9e5, 1e7,
                                  np.inf, np.nan])])
    x = np.ascontiguousarray(x)

    def run(func):
        out = np.empty_like(x)
        assert lib.tvd_debug_math(0, func, x.size, _lib.ptr_f64(x), _lib.ptr_f64(out)) == 0
        return out
    dd_s, dd_c = run(0), run(1)
    for func, want in ((3, dd_s), (4, dd_c), (5, dd_s), (6, dd_c)):
        got = run(func)
        same = (got == want) | (np.isnan(got) & np.isnan(want))
        assert np.all(same), (func, int(np.sum(~same)), x[~same][:5])
    assert np.signbit(run(3)[x.size - 13])            # sin(-0.0) = -0.0
    idx = rng.choice(1500000, 3000, replace=False)
    ms = np.array([float(mpmath.sin(mpmath.mpf(float(v)))) for v in x[idx]])
    mc = np.array([float(mpmath.cos(mpmath.mpf(float(v)))) for v in x[idx]])
    assert np.array_equal(run(3)[idx], ms) and np.array_equal(run(4)[idx], mc)


def test_table_exp_equals_double_double():
    """K1/K4's exp (tvd_crtrig.cuh: table + Ziv's rounding test) returns the bits of the
    double-double version, i.e. the correctly rounded value."""
    import mpmath
    mpmath.mp.dps = 60
    lib = _lib.load()
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(-1, 1, 300000), rng.uniform(-40, 5, 400000),
                        rng.uniform(-690, 690, 200000), rng.uniform(-1e-6, 1e-6, 50000),
                        np.array([0.0, -0.0, 1e-300, 1.0, -1.0, 699.9, -699.9, 700.0, 710.0, -745.0,
                                  -800.0, np.inf, -np.inf, np.nan])])
    x = np.ascontiguousarray(x)

    def run(func):
        out = np.empty_like(x)
        assert lib.tvd_debug_math(0, func, x.size, _lib.ptr_f64(x), _lib.ptr_f64(out)) == 0
        return out
    dd, cr = run(2), run(7)
    same = (dd == cr) | (np.isnan(dd) & np.isnan(cr))
    assert np.all(same), (int(np.sum(~same)), x[~same][:5])
    idx = rng.choice(900000, 3000, replace=False)
    want = np.array([float(mpmath.exp(mpmath.mpf(float(v)))) for v in x[idx]])
    assert np.array_equal(cr[idx], want)


def test_k1_tie_breaking_with_whole_periods(oracle_lib):
    """A sine with a whole number of periods over the tesseroid has mathematically equal maxima
    of |rho_n - line|; the split list is then decided by the last bit of sin()."""
    from tesseroid_variable_density_b200.model import FlatModel
    rng = np.random.default_rng(7)
    n = 300
    tops = rng.uniform(-1000, 0, n)
    bots = tops - rng.uniform(5000, 50000, n)
    bounds = np.zeros((n, 6))
    bounds[:, 1] = 1
    bounds[:, 3] = 1
    bounds[:, 4], bounds[:, 5] = tops, bots
    kinds = np.full(n, D.SINE, dtype=np.int32)
    params = np.zeros((n, 5))
    params[:, 0] = 1650.
    params[:, 1] = 2 * np.pi * rng.integers(1, 5, n) / (tops - bots)
    params[:, 2] = 1650.
    for delta in (0.1, 1e-2):
        pm = tesseroid.PreparedModel(FlatModel(bounds, kinds, params), delta=delta)
        t, b, parent = pm.pieces()
        pm.close()
        pos, bad = 0, 0
        for i in range(n):
            ot, ob = oracle_lib.density_discretize(tops[i], bots[i], kinds[i], params[i], delta)
            k = int(np.sum(parent == i))
            if k != len(ot) or not (np.array_equal(t[pos:pos + k], ot) and np.array_equal(b[pos:pos + k], ob)):
                bad += 1
            pos += k
        assert bad == 0, "%d of %d tesseroids split differently at delta=%g" % (bad, n, delta)


def test_plain_c_client_of_the_abi(tmp_path, oracle_lib):
    """examples/c_abi_demo.c: the C ABI used from plain C (gcc), no Python in the loop; its
    printed numbers must agree with the oracle."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "tesseroid_variable_density_b200", "csrc")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "c_abi_demo.c"), "-o", exe, "-L", csrc,
                           "-ltvd", "-lm", "-Wl,-rpath," + csrc])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    gz = float(re.search(r"gz at the centre point\s+(\S+)", out).group(1))
    pot = float(re.search(r"potential at the centre\s+(\S+)", out).group(1))
    mesh = TesseroidMesh((-1.5, 1.5, -1.5, 1.5, 0, -35000), (1, 3, 3))
    mesh.addprop("density", [D.ExponentialDensity.between(2670, 3300, 0., -35000., 3)] * 9)
    lats, lons, heights = gridder.regular((-2, 2, -2, 2), (5, 5), z=10e3)
    want_gz = cases.OracleEngine(oracle_lib)("gz", lons, lats, heights, mesh, 2.5, 0.1)[12]
    want_pot = cases.OracleEngine(oracle_lib)("potential", lons, lats, heights, mesh, 1, 0.1)[12]
    assert abs(gz - want_gz) <= 1e-5 and abs(pot - want_pot) <= 1e-5      # printed with 6 decimals
