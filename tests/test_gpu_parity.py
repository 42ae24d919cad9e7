"""
Parity of the CUDA path (through the C ABI) with the oracle on the same inputs.

Bars:
  * GLQ evaluations per (tesseroid, point) and all node/leaf/split counters: identical;
  * field values: <= 1e-9 relative at heights >= 2 km for potential, gx, gy, gz;
    the gradient tensor is l^-5-conditioned and only pinned by the analytical shell, it gets
    1e-6 at 2 km and 1e-10 at 260 km against the oracle;
  * z = 0 grids over thin shells: the reference's own output moves by up to 6.6e-6 under 1-ulp
    perturbations of its inputs, so those get a conditioning-scaled 1e-4.
"Relative" is relative to the largest magnitude over the grid of the field's family
(gx, gy -> gz; tensor components -> gzz) because gx, gy and the off-diagonal tensor components of
a closed shell vanish by symmetry and have no scale of their own.
"""
import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import density as D
from tesseroid_variable_density_b200 import gridder
from tesseroid_variable_density_b200.mesher import TesseroidMesh, TesseroidModel
from tesseroid_variable_density_b200.model import flatten_model
from tesseroid_variable_density_b200.tesseroid import PreparedModel, _convert_coords, forward_raw

pytestmark = pytest.mark.gpu

FAMILY = {"potential": "potential", "gx": "gz", "gy": "gz", "gz": "gz"}
for _f in cases.ALL_FIELDS[4:]:
    FAMILY[_f] = "gzz"
COUNT_KEYS = ("nodes", "leaves", "splits", "too_small", "pieces", "nonroot_nodes",
              "nonroot_leaves")


def run_both(oracle, mesh, grid, fields, delta, ratios=None, counts=True):
    """{field: (gpu, oracle, rel, counts_equal, stats_equal)} in kernel units."""
    lats, lons, heights = grid
    flat = flatten_model(mesh)
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=delta)
    out, scale = {}, {}
    try:
        for field in fields:
            ratio = (ratios or cases.DEFAULT_RATIO)[field] if not np.isscalar(ratios) else ratios
            o = oracle.forward(field, flat.bounds, flat.kind, flat.params, delta, ratio, *cc,
                               leaf_counts=counts, n_threads=8)
            g = forward_raw(field, *cc, pm, ratio, leaf_counts=counts, return_stats=True)
            out[field] = (g, o)
            scale[field] = np.max(np.abs(o[0]))
    finally:
        pm.close()
    res = {}
    for field, (g, o) in out.items():
        fam = FAMILY[field]
        s = max(scale[field], scale.get(fam, 0.0))
        rel = np.max(np.abs(g[0] - o[0])) / s if s > 0 else float(np.max(np.abs(g[0] - o[0])))
        ceq = np.array_equal(g[2], o[2]) if counts else True
        seq = all(g[1][k] == o[1][k] for k in COUNT_KEYS)
        res[field] = (g[0], o[0], rel, ceq, seq)
    return res


SHELLS = [
    ("exp30_1km", lambda: cases.exponential_case(1e3, 30)[0], 0.1),
    ("exp1_100km", lambda: cases.exponential_case(1e5, 1)[0], 0.1),
    ("sine5_100km", lambda: cases.sine_case(1e5, 5)[0], 1e-2),
    ("linear_10km", lambda: cases.linear_case(1e4)[0], None),
]


@pytest.mark.parametrize("name,mk,delta", SHELLS, ids=[s[0] for s in SHELLS])
@pytest.mark.parametrize("grid", ["global2km", "equator2km", "pole2km", "260km"])
def test_shell_gravity_fields(oracle_lib, name, mk, delta, grid):
    res = run_both(oracle_lib, mk(), cases.GRIDS[grid](), ["gz", "potential", "gx", "gy"], delta)
    for field, (_, _, rel, ceq, seq) in res.items():
        assert ceq, "%s: GLQ counts differ" % field
        assert seq, "%s: counters differ" % field
        assert rel <= 1e-9, "%s: rel %.3e" % (field, rel)       # tolerance: 1e-9 (north_star)


@pytest.mark.parametrize("grid,tol", [("global2km", 1e-6), ("260km", 1e-10)])
def test_shell_gradient_tensor(oracle_lib, grid, tol):
    mesh = cases.exponential_case(1e3, 30)[0]
    fields = ["gzz", "gxx", "gxy", "gxz", "gyy", "gyz"]
    res = run_both(oracle_lib, mesh, cases.GRIDS[grid](), fields, 0.1)
    for field, (_, _, rel, ceq, seq) in res.items():
        assert ceq and seq, field
        assert rel <= tol, "%s: rel %.3e" % (field, rel)


@pytest.mark.parametrize("grid", ["global", "pole", "equator"])
def test_shell_at_zero_height_conditioning_scaled(oracle_lib, grid):
    mesh = cases.exponential_case(1e3, 30)[0]
    res = run_both(oracle_lib, mesh, cases.GRIDS[grid](), ["gz", "potential"], 0.1)
    for field, (_, _, rel, ceq, seq) in res.items():
        assert ceq and seq, field
        assert rel <= 1e-4, "%s: rel %.3e" % (field, rel)       # conditioning-scaled (see above)


@pytest.mark.parametrize("ratio", [1, 1.5, 2, 2.5, 3, 4, 5])
def test_ratio_sweep_counts_and_values(oracle_lib, ratio):
    # BASELINE config 2: sweep of the distance-size ratio 1..5
    mesh = cases.exponential_case(1e4, 10)[0]
    lats, lons, heights = cases.GRIDS["global2km"]()
    flat = flatten_model(mesh)
    cc = _convert_coords(lons, lats, heights)
    pm = PreparedModel(mesh, delta=0.1)
    try:
        for field in ("gz", "potential"):
            o = oracle_lib.forward(field, flat.bounds, flat.kind, flat.params, 0.1, ratio, *cc,
                                   leaf_counts=True, n_threads=8)
            g = forward_raw(field, *cc, pm, ratio, leaf_counts=True)
            assert np.array_equal(g[2], o[2])
            assert all(g[1][k] == o[1][k] for k in COUNT_KEYS)
            assert np.max(np.abs(g[0] - o[0])) <= 1e-9 * np.max(np.abs(o[0]))
    finally:
        pm.close()


@pytest.mark.parametrize("delta", [0.1, 0.2, 0.5, 1.0, 1e-3])
@pytest.mark.parametrize("kind", ["exp", "sine"])
def test_delta_sweep(oracle_lib, kind, delta):
    # BASELINE config 2: sweep of the delta ratio 0.1..1 (plus a fine one)
    mesh = cases.exponential_case(1e4, 30)[0] if kind == "exp" else cases.sine_case(1e4, 2)[0]
    res = run_both(oracle_lib, mesh, cases.GRIDS["global2km"](), ["gz"], delta)
    _, _, rel, ceq, seq = res["gz"]
    assert ceq and seq and rel <= 1e-9, rel


def regional_model(nlat=40, nlon=50, seed=0, kind="exp"):
    rng = np.random.default_rng(seed)
    top = np.full(nlat * nlon, 845.0)
    bottom = top - rng.uniform(200, 6000, nlat * nlon)
    tm = TesseroidModel((-40., -38.05, 288., 290.45), top, bottom, (nlat, nlon))
    if kind == "exp":
        dens = [D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)] * tm.size
    elif kind == "mixed":
        # notebook cell 12: constants, a linear function and exponentials in one model
        e = D.ExponentialDensity.between(-412, -275, 845.0, -5155.0, 3)
        dens = [e] * tm.size
        dens[0] = 500
        dens[1] = D.LinearDensity(-0.01, 0., 0.)
        dens[2] = D.SineDensity(100., 2 * np.pi / 3000., -300.)
    tm.addprop("density", dens)
    return tm


@pytest.mark.parametrize("z", [2000., 10e3])
@pytest.mark.parametrize("kind", ["exp", "mixed"])
def test_regional_model_all_gravity_fields(oracle_lib, z, kind):
    tm = regional_model(kind=kind)
    grid = gridder.regular((-39.9, -38.1, 288.1, 290.4), (30, 33), z=z)
    res = run_both(oracle_lib, tm, grid, ["gz", "potential", "gx", "gy"], 0.1, counts=False)
    for field, (g, o, _, _, seq) in res.items():
        assert seq, field
        own = np.max(np.abs(g - o)) / np.max(np.abs(o))     # gx, gy have their own scale here
        assert own <= 1e-9, "%s: rel %.3e" % (field, own)


def test_regional_tensor_all_components(oracle_lib):
    tm = regional_model(nlat=16, nlon=16)
    grid = gridder.regular((-39.9, -39.3, 288.1, 288.7), (12, 12), z=10e3)
    res = run_both(oracle_lib, tm, grid, list(cases.ALL_FIELDS[4:]), 0.1, counts=True)
    for field, (g, o, _, ceq, seq) in res.items():
        assert ceq and seq, field
        own = np.max(np.abs(g - o)) / np.max(np.abs(o))
        assert own <= 1e-8, "%s: rel %.3e" % (field, own)


def test_neuquen_sample_vs_oracle_and_stored(oracle_lib):
    """BASELINE config 3 on a sample of its grid: oracle parity <= 1e-9 and the reference's
    stored result (tests/golden/reference_results/neuquen/exponential-gz.npz) <= 1e-9."""
    import os
    stored = np.load(os.path.join(cases.REF_RESULTS, "neuquen", "exponential-gz.npz"))
    basin, top, bottom = cases.neuquen_model()
    basin.addprop("density", [cases.neuquen_densities(top, bottom)["exponential"]] * basin.size)
    idx = np.linspace(0, stored["lon"].size - 1, 300).astype(int)
    lons, lats, heights = stored["lon"][idx], stored["lat"][idx], stored["height"][idx]
    res = run_both(oracle_lib, basin, (lats, lons, heights), ["gz"], 1e-2, counts=False)
    g, o, rel, _, seq = res["gz"]
    assert seq and rel <= 1e-9
    got = g * cases.UNIT["gz"]
    assert np.max(np.abs(got - stored["result"][idx])) <= 1e-9 * np.max(np.abs(stored["result"]))


def test_degenerate_and_tiny_inputs(oracle_lib):
    # zero-thickness cells are still evaluated (neuquen_basin.py:69-72) and add exact zeros;
    # one point, one tesseroid; a point right on top of a tesseroid (deep subdivision)
    t = [TesseroidMesh((10, 11, 20, 21, 0, 0), (1, 1, 1)), TesseroidMesh((10, 11, 20, 21, 100., -900.), (1, 1, 1))]
    for m in t:
        m.addprop("density", [D.ExponentialDensity.between(2670, 3300, 100., -900., 5)])
    grid = (np.array([20.5]), np.array([10.5]), np.array([150.]))
    for mesh in t:
        res = run_both(oracle_lib, mesh, grid, ["gz", "potential"], 0.1)
        for field, (g, o, rel, ceq, seq) in res.items():
            assert ceq and seq
            if np.max(np.abs(o)) > 0:
                assert rel <= 1e-7, rel       # 50 m above the cell: conditioning-limited
            else:
                assert np.all(g == 0)


def test_notebook_cases(oracle_lib):
    """The two models of the reference's notebook (code/notebooks/tesseroid_density_example.ipynb,
    cells 4-16): one tesseroid with an exponential density on a 51 x 51 grid at 260 km, and a
    3 x 3 mesh mixing a constant, a linear and exponential densities at 50 km; all four fields."""
    from tesseroid_variable_density_b200.mesher import Tesseroid
    top, bottom = 0, -35000
    rho0, rho1, b = 2670, 3300, 1e5
    a = (rho1 - rho0) / (np.exp((abs(top - bottom)) / b) - 1)
    c = rho0 - a
    dens = D.ExponentialDensity(a, -1.0, 0.0, b, c)          # a*np.exp(-height/b) + c  (cell 6)
    h = np.linspace(bottom, top, 11)
    assert np.array_equal(dens(h), a * np.exp(-h / b) + c)
    model = [Tesseroid(-0.1, 0.1, -0.1, 0.1, top, bottom, props={"density": dens})]
    grid = gridder.regular((-5, 5, -5, 5), (51, 51), z=260000)
    res = run_both(oracle_lib, model, grid, ["potential", "gx", "gy", "gz"], 1e-2, counts=False)
    for field, (g, o, _, _, seq) in res.items():
        assert seq and np.max(np.abs(g - o)) <= 1e-9 * np.max(np.abs(o)), field
    mesh = TesseroidMesh((-1, 1, -1, 1, top, bottom), (1, 3, 3))
    dl = [dens] * mesh.size
    dl[0] = 500
    dl[1] = D.LinearDensity(-0.01, 0.0, 0.0)                 # -0.01*height  (cell 12)
    mesh.addprop("density", dl)
    grid = gridder.regular((-5, 5, -5, 5), (51, 51), z=50000)
    res = run_both(oracle_lib, mesh, grid, ["potential", "gx", "gy", "gz"], 1e-2, counts=True)
    for field, (g, o, _, ceq, seq) in res.items():
        assert ceq and seq and np.max(np.abs(g - o)) <= 1e-9 * np.max(np.abs(o)), field


@pytest.mark.parametrize("size_deg,safe_eps", [(1e-3, 1e-4), (1e-4, 1e-2), (1e-5, 0.5)])
def test_tiny_cells_at_the_ratio_boundary(oracle_lib, monkeypatch, size_deg, safe_eps):
    """Cells of 0.001 ... 0.00001 degree at distances d = D*L*(1 +- eps) from the point.
    L = rtop*acos(arg) with 1 - arg ~ 1e-10 ... 1e-14 for such cells: a 1-ulp difference in arg
    moves L by 7e-7 ... 7e-3, more than a fixed relative guard on the far-field test covers
    (divisions, _tesseroid.pyx:129-146 with the sizes of :104-123).
      * at |eps| >= safe_eps (10x the conditioning of L) the subdivision counts must be the
        oracle's;
      * at every eps, including 1e-9, the far-field shortcut must agree with the literal walk of
        the same pair (TVD_DEBUG_NO_FAR=1 sends every pair through the walk)."""
    from tesseroid_variable_density_b200.mesher import Tesseroid
    R, D_ratio, d2r = 6378137.0, 2.5, np.pi / 180.
    w, s = 10.0, 20.0
    e, n = w + size_deg, s + size_deg
    top = 0.0
    bottom = -max(1.0, 1.1e5 * size_deg)
    model = [Tesseroid(w, e, s, n, top, bottom, props={"density": 1000.})]
    rt, rtop = 0.5 * (top + bottom) + R, top + R
    latt = d2r * 0.5 * (s + n)
    Llon = rtop * np.arccos(np.sin(latt) ** 2 + np.cos(latt) ** 2 * np.cos(d2r * (e - w)))
    Llat = rtop * np.arccos(np.sin(d2r * n) * np.sin(d2r * s) + np.cos(d2r * n) * np.cos(d2r * s))
    eps = np.array([1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 0.1, 0.5])
    eps = np.concatenate([-eps[::-1], eps])
    dist = np.concatenate([D_ratio * Llon * (1 + eps), D_ratio * Llat * (1 + eps)])
    all_eps = np.concatenate([eps, eps])
    lons = np.full(dist.size, 0.5 * (w + e))
    lats = np.full(dist.size, 0.5 * (s + n))
    heights = (rt - R) + dist
    flat = flatten_model(model)
    cc = _convert_coords(lons, lats, heights)
    o = oracle_lib.forward("gz", flat.bounds, flat.kind, flat.params, 0.1, D_ratio, *cc,
                           leaf_counts=True, n_threads=1)
    pm = PreparedModel(model, delta=0.1)
    g = forward_raw("gz", *cc, pm, D_ratio, leaf_counts=True, return_stats=True)
    pm.close()
    monkeypatch.setenv("TVD_DEBUG_NO_FAR", "1")
    pm = PreparedModel(model, delta=0.1)
    lit = forward_raw("gz", *cc, pm, D_ratio, leaf_counts=True, return_stats=True)
    pm.close()
    monkeypatch.delenv("TVD_DEBUG_NO_FAR")
    assert lit[1]["near_pairs"] == dist.size
    assert np.array_equal(g[2], lit[2])                 # shortcut == literal walk, every eps
    assert np.allclose(g[0], lit[0], rtol=1e-9, atol=0)      # sweep vs walk arithmetic, 1e-9 bar
    safe = np.abs(all_eps) >= safe_eps
    assert np.array_equal(g[2][0][safe], o[2][0][safe]), (g[2][0][safe], o[2][0][safe])
    assert np.max(np.abs(g[0][safe] - o[0][safe])) <= 1e-9 * np.max(np.abs(o[0][safe]))
    assert len(set(o[2][0])) >= 2                       # the boundary is actually crossed


def test_c5_geometry_fused_fields(oracle_lib):
    """BASELINE.json configs[4] (SURVEY.md 8d "C5") at a size the oracle finishes: a global
    60 x 120 crust of per-cell sinusoidal densities (two periods over each cell's own
    thickness, ~5 radial pieces per cell), potential + gz + gzz at 260 km in one fused sweep
    (ratios 1 / 2.5 / 8), against the oracle field by field: counters identical, values
    <= 1e-9 (potential, gz) and <= 1e-6 (gzz, l^-5-conditioned) relative."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scripts"))
    from c5_full import c5_model, c5_grid, NAMES, RATIOS
    from tesseroid_variable_density_b200.tesseroid import forward_raw_fused
    flat = c5_model(60, 120)
    lon, lat, h = c5_grid(12, 24, 260e3)
    cc = _convert_coords(lon, lat, h)
    pm = PreparedModel(flat, delta=0.1)
    try:
        assert pm.n_pieces > 4 * flat.size
        out, st = forward_raw_fused(NAMES, *cc, pm, RATIOS, return_stats=True)
        for f, r, tol in zip(NAMES, RATIOS, (1e-9, 1e-9, 1e-6)):
            o, so = oracle_lib.forward(f, flat.bounds, flat.kind, flat.params, 0.1, r, *cc,
                                       n_threads=8)
            assert all(st[f][k] == so[k] for k in COUNT_KEYS), f
            rel = np.max(np.abs(out[f] - o)) / np.max(np.abs(o))
            assert rel <= tol, "%s: rel %.3e" % (f, rel)
            single, _ = forward_raw(f, *cc, pm, r, return_stats=True)
            assert np.max(np.abs(single - out[f])) <= 1e-13 * np.max(np.abs(o))
    finally:
        pm.close()
