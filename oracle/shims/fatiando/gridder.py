"""
``gridder.regular`` of Fatiando 0.5: the first coordinate of ``area`` varies slowest.
Pinned by the golden replay in tests/test_oracle_ref.py.
"""
import numpy as np


def regular(area, shape, z=None):
    nx, ny = shape
    x1, x2, y1, y2 = area
    xs = np.linspace(x1, x2, nx)
    ys = np.linspace(y1, y2, ny)
    arrays = list(np.meshgrid(ys, xs))[::-1]
    if z is not None:
        arrays.append(z*np.ones(nx*ny, dtype=float))
    return [i.ravel() for i in arrays]
