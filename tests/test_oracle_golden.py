"""
Pins the oracle (oracle/tess_oracle.c) against the reference's own stored golden results
(tests/golden/reference_results = code/scripts/results of the reference): the `differences`
arrays of the shell-vs-analytical sweeps must be reproduced to the last bit (the reference
build replays them bit-exactly), the Neuquen grids to 1e-10.
"""
import os

import numpy as np
import pytest

import cases

# (family, file, number of sweep points replayed) -- a bounded subset so the CPU suite stays
# within minutes; the GPU suite replays a different, overlapping subset.
GOLDENS = [
    ("linear", "potential-global-1000.npz", 10),
    ("linear", "gz-global-1000.npz", 6),
    ("linear", "gz-260km-100000.npz", 10),
    ("linear", "potential-pole-10000.npz", 10),
    ("linear", "gz-equator-1000000.npz", 5),
    ("exponential", "potential-global-1000-30.npz", 9),
    ("exponential", "gz-260km-1000000-100.npz", 9),
    ("exponential", "gz-global-1000-1.npz", 3),
    ("exponential", "potential-equator-100000-5.npz", 9),
    ("sine", "potential-260km-1000-10.npz", 5),
    ("sine", "potential-260km-100000-5.npz", 5),
    ("sine", "gz-260km-10000-2.npz", 5),
    ("grid-search", "potential-global-1000-30.npz", 8),
]


@pytest.mark.parametrize("family,fname,nsweep", GOLDENS)
def test_oracle_reproduces_stored_differences(oracle_lib, family, fname, nsweep):
    engine = cases.OracleEngine(oracle_lib)
    got, want = cases.replay_golden(engine, family, fname, max_sweep=nsweep)
    # bit-exact wherever glibc and NumPy agree on exp/sin of the density (always, in practice)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    assert np.mean(got == want) >= 0.8


def _neuquen(oracle_lib, kind, field, npoints, rtol):
    stored = np.load(os.path.join(cases.REF_RESULTS, "neuquen", "%s-%s.npz" % (kind, field)))
    basin, top, bottom = cases.neuquen_model()
    dens = cases.neuquen_densities(top, bottom)[kind]
    basin.addprop("density", [dens] * basin.size)
    idx = np.linspace(0, stored["lon"].size - 1, npoints).astype(int)
    engine = cases.OracleEngine(oracle_lib)
    got = engine(field, stored["lon"][idx], stored["lat"][idx], stored["height"][idx], basin,
                 cases.DEFAULT_RATIO[field], 1e-2)
    want = stored["result"][idx]
    assert np.max(np.abs(got - want)) <= rtol * np.max(np.abs(stored["result"]))


def test_oracle_neuquen_homogeneous_gz(oracle_lib):
    _neuquen(oracle_lib, "homogeneous", "gz", 400, 1e-10)


@pytest.mark.parametrize("kind,field", [("exponential", "gz"), ("exponential", "potential"),
                                        ("linear", "gz")])
def test_oracle_neuquen_variable_density(oracle_lib, kind, field):
    _neuquen(oracle_lib, kind, field, 200, 1e-10)
