#!/usr/bin/env python
"""
Spherical-shell validation of the paper (reference: code/scripts/exponential_density.py): a
shell of 72 tesseroids with exponential density against the analytical solution, for all ten
fields, at the paper's default accuracy controls (D = 1 / 2.5 / 8, delta = 0.1).

    python examples/shell_validation.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tesseroid_variable_density_b200 import (ExponentialDensity, TesseroidMesh, analytical,  # noqa: E402
                                             gridder, tesseroid)

thickness, b_factor = 1e4, 10
mesh = TesseroidMesh((0, 360, -90, 90, 0, -thickness), (1, 6, 12))
A = (3300 - 2670) / (1 - np.exp(-b_factor))
C = 3300 - A
mesh.addprop("density", [ExponentialDensity(A, -b_factor, -thickness, thickness, C)] * mesh.size)
lats, lons, heights = gridder.regular((-90, 90, 0, 360), (19, 13), z=260e3)
exact = analytical.shell_exponential_density(heights[0], 0, -thickness, A, b_factor, C)
names = ["potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz"]
vector = tesseroid.fields(["potential", "gx", "gy", "gz"], lons, lats, heights, mesh, delta=0.1)
tensor = tesseroid.fields(names[4:], lons, lats, heights, mesh, delta=0.1)
for f in names:
    got = vector[f] if f in vector else tensor[f]
    if exact[f] != 0:
        print("%-9s analytical % .6e   max difference %.4f %%" % (
            f, exact[f], 100 * np.max(np.abs((got - exact[f]) / exact[f]))))
    else:
        print("%-9s analytical 0 (symmetry)   max |value| %.3e" % (f, np.max(np.abs(got))))
