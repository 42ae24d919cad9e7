"""
``regular``: the point-grid generator the reference scripts take from Fatiando 0.5
(``fatiando.gridder.regular``; e.g. code/scripts/linear_density.py:54-57).  The first pair of
``area`` varies slowest, so the scripts unpack ``lats, lons, heights = regular(...)``.
"""
import numpy as np


def regular(area, shape, z=None):
    nx, ny = shape
    x1, x2, y1, y2 = area
    xs = np.linspace(x1, x2, nx)
    ys = np.linspace(y1, y2, ny)
    arrays = list(np.meshgrid(ys, xs))[::-1]
    if z is not None:
        arrays.append(z * np.ones(nx * ny, dtype=float))
    return [i.ravel() for i in arrays]
