#!/usr/bin/env python
"""
The Neuquen Basin application of the paper (reference: code/scripts/neuquen_basin.py) with the
B200 library: sediment thickness -> tesseroid model -> potential and gz at 10 km height for a
homogeneous, a linear and an exponential density contrast.  Compares with the reference's stored
results (tests/golden/reference_results/neuquen) when they are available.  No plotting.

    python examples/neuquen_basin.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tesseroid_variable_density_b200 import (ExponentialDensity, LinearDensity, datasets,  # noqa: E402
                                             gridder, tesseroid)

GOLDEN = os.path.join(ROOT, "tests", "golden")      # only for the optional comparison below

# data/sediment_thickness.dat (117 x 91 nodes of 0.05 deg; NaN outside the basin) -> sediment
# layer with a constant top (median topography over the basin, 845 m in the paper)
sediments = datasets.load_sediment_thickness(os.path.join(ROOT, "data", "sediment_thickness.dat"))
basin, layer_top, layer_bottom = datasets.neuquen_basin_model(sediments)
print("model: %d tesseroids, sediments between %.1f and %.1f m" % (basin.size, layer_top, layer_bottom))

# computation grid at 10 km
glat, glon, gheight = gridder.regular((-40.8, -33., 287, 295.), (79, 81), z=10e3)

density_top, density_bottom = -412, -275
densities = {
    "homogeneous": (density_top + density_bottom) / 2,
    "linear": LinearDensity.between(density_top, density_bottom, layer_top, layer_bottom),
    "exponential": ExponentialDensity.between(density_top, density_bottom, layer_top, layer_bottom, 3),
}
for name, dens in densities.items():
    basin.addprop("density", [dens] * basin.size)
    t0 = time.perf_counter()
    out = tesseroid.fields(["potential", "gz"], glon, glat, gheight, basin)
    dt = time.perf_counter() - t0
    line = "%-12s potential+gz on %d points: %.3f s; gz range %.3f .. %.3f mGal" % (
        name, glon.size, dt, out["gz"].min(), out["gz"].max())
    stored = os.path.join(GOLDEN, "reference_results", "neuquen", "%s-gz.npz" % name)
    if os.path.isfile(stored):
        want = np.load(stored)["result"]
        line += "; max |diff| to the reference's stored result %.1e mGal" % np.max(np.abs(out["gz"] - want))
    print(line)
