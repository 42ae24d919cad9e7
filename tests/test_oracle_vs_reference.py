"""
Pins the oracle against the UNMODIFIED reference compiled into oracle/_ref
(oracle/build_ref.py): field values bit for bit, the radial split lists exactly, and the
per-(tesseroid, point) GLQ counts through the counting-callable trick.
"""
import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import density as D
from tesseroid_variable_density_b200 import gridder
from tesseroid_variable_density_b200.mesher import Tesseroid, TesseroidMesh
from tesseroid_variable_density_b200.model import flatten_model
from tesseroid_variable_density_b200.tesseroid import _convert_coords


def _regional_mesh(dens):
    mesh = TesseroidMesh((-2, 2, -1.5, 1.5, 500., -4000.), (1, 6, 8))
    mesh.addprop("density", [dens] * mesh.size)
    return mesh


DENSITIES = {
    "exp": D.ExponentialDensity.between(-412, -275, 500., -4000., 3),
    "sine": D.SineDensity(1650., 2*np.pi*2/4500., 1650.),
    "linear": D.LinearDensity.between(2670, 3300, 500., -4000.),
    "const": 2670.0,
}


@pytest.mark.parametrize("field", ["potential", "gx", "gy", "gz"])
@pytest.mark.parametrize("dname", sorted(DENSITIES))
def test_fields_bit_identical_to_reference(oracle_lib, reference, field, dname):
    mesh = _regional_mesh(DENSITIES[dname])
    lats, lons, heights = gridder.regular((-3, 3, -3, 3), (9, 9), z=2000.)
    ratio = cases.DEFAULT_RATIO[field]
    want = cases.ReferenceEngine(reference)(field, lons, lats, heights, mesh, ratio, 0.1)
    got = cases.OracleEngine(oracle_lib, n_threads=1)(field, lons, lats, heights, mesh, ratio, 0.1)
    assert np.array_equal(got, want)


def test_delta_none_and_njobs(oracle_lib, reference):
    mesh = _regional_mesh(DENSITIES["exp"])
    lats, lons, heights = gridder.regular((-3, 3, -3, 3), (7, 5), z=10e3)
    want = reference.gz(lons, lats, heights, mesh, delta=None)
    got1 = cases.OracleEngine(oracle_lib, n_threads=1)("gz", lons, lats, heights, mesh, 2.5, None)
    got3 = cases.OracleEngine(oracle_lib, n_threads=3)("gz", lons, lats, heights, mesh, 2.5, None)
    assert np.array_equal(got1, want)
    assert np.array_equal(got3, want)      # point chunking does not change any value


class CountingDensity(object):
    """Counts scalar calls: the reference calls the density 8 times per integrated cell
    (_tesseroid.pyx:515) and only with arrays during the radial discretisation."""

    def __init__(self, inner):
        self.inner = inner
        self.calls = 0

    def __call__(self, h):
        if np.ndim(h) == 0:
            self.calls += 1
        return self.inner(h)


def test_leaf_counts_match_reference(oracle_lib, reference):
    dens = DENSITIES["exp"]
    mesh = _regional_mesh(dens)
    flat = flatten_model(mesh)
    lats, lons, heights = gridder.regular((-2.5, 2.5, -2.5, 2.5), (4, 4), z=2000.)
    cc = _convert_coords(lons, lats, heights)
    _, stats, counts = oracle_lib.forward("gz", flat.bounds, flat.kind, flat.params, 0.1, 2.5,
                                          *cc, leaf_counts=True)
    assert counts.sum() == stats["leaves"]
    # per (tesseroid, point) through the reference, one pair at a time
    for t in (0, 13, 27, 47):
        b = flat.bounds[t]
        for p in (0, 5, 10, 15):
            cd = CountingDensity(dens)
            tess = [Tesseroid(*b, props={"density": cd})]
            reference.gz(lons[p:p+1], lats[p:p+1], heights[p:p+1], tess, delta=0.1)
            assert cd.calls % 8 == 0
            assert cd.calls // 8 == counts[t, p]


# Known answers of the density-based discretisation (the reference's
# number-of-tesseroids.py:56-80,103-130), bounds [-10, 10, -10, 10, 0, -1000], delta = 0.1
def _exp01(b):
    A = (1 - 0) / (1 - np.exp(-b))
    return D.ExponentialDensity(A, -b, -1000., 1000., 1 - A)


def _sine01(b):
    return D.SineDensity(0.5, 2*np.pi*b/1000., 0.5)


KNOWN = [
    (_exp01(1), [(0, -540), (-540, -1000)]),
    (_exp01(2), [(0, -580), (-580, -1000)]),
    (_exp01(5), [(0, -680), (-680, -1000)]),
    (_exp01(10), [(0, -770), (-770, -1000)]),
    (_exp01(30), [(0, -890), (-890, -1000)]),
    (_exp01(100), [(0, -950), (-950, -1000)]),
    (_sine01(1), [(0, -250), (-250, -715), (-715, -1000)]),
]


@pytest.mark.parametrize("dens,pieces", KNOWN)
def test_discretization_known_answers(oracle_lib, dens, pieces):
    tops, bots = oracle_lib.density_discretize(0., -1000., dens.kind, dens.params, 0.1)
    got = [(round(t, 6), round(b, 6)) for t, b in zip(tops, bots)]
    assert got == [(float(t), float(b)) for t, b in pieces]


@pytest.mark.parametrize("b,npieces", [(2, 5), (5, 11), (10, 20)])
def test_discretization_sine_piece_counts(oracle_lib, b, npieces):
    dens = _sine01(b)
    tops, bots = oracle_lib.density_discretize(0., -1000., dens.kind, dens.params, 0.1)
    assert len(tops) == npieces


@pytest.mark.parametrize("delta", [1e-3, 3.16e-3, 1e-2, 0.1, 1.0])
@pytest.mark.parametrize("dname", ["exp", "sine", "linear"])
def test_discretization_identical_to_reference(oracle_lib, reference, dname, delta):
    dens = DENSITIES[dname]
    bounds = [-1., 1., -1., 1., 500., -4000.]
    want = reference._density_based_discretization(np.array(bounds), dens, delta)
    tops, bots = oracle_lib.density_discretize(500., -4000., dens.kind, dens.params, delta)
    assert len(want) == len(tops)
    for w, t, b in zip(want, tops, bots):
        assert w[4] == t and w[5] == b


def test_shell_exp_piece_counts(oracle_lib):
    # exp shell 3300/2670, T = 1 km, b = 30
    _, dens, _ = cases.exponential_case(1e3, 30)
    for delta, n in [(1e-3, 7), (3.16e-3, 6), (1e-2, 4), (3.16e-2, 3), (0.1, 2), (0.316, 2),
                     (1.0, 1)]:
        tops, _ = oracle_lib.density_discretize(0., -1e3, dens.kind, dens.params, delta)
        assert len(tops) == n
