"""
Input data of the Neuquén Basin application (BASELINE.json configs[2]).

* :func:`load_sediment_thickness` reads ``sediment_thickness.dat`` -- three ASCII columns
  latitude, longitude, thickness (m), ``nan`` outside the basin -- exactly as the reference does
  (``numpy.loadtxt(..., unpack=True)``, data/README.md:25-37 and
  code/scripts/neuquen_basin.py:33-47).
* :func:`neuquen_basin_model` builds the sediment layer of code/scripts/neuquen_basin.py:52-72:
  a :class:`~.mesher.TesseroidModel` of 117 x 91 cells whose top is one constant (the median
  topography over the basin) and whose bottom is ``top - thickness``; cells outside the basin
  become degenerate ``top = bottom = 0`` tesseroids that are still evaluated, as in the
  reference.
"""
from __future__ import annotations

import os

import numpy as np

from .mesher import TesseroidModel

#: median topography over the basin.  The reference derives it from data/topography.npy, which
#: is missing upstream (.MISSING_LARGE_BLOBS); this value reproduces its stored results to
#: <= 3e-11.
NEUQUEN_TOP = 845.0761706691
NEUQUEN_SHAPE = (117, 91)

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_THICKNESS_FILE = os.path.join(os.path.dirname(_HERE), "data", "sediment_thickness.dat")


def load_sediment_thickness(path=None, shape=NEUQUEN_SHAPE):
    """``{'lat', 'lon', 'thickness', 'shape', 'area'}`` of ``sediment_thickness.dat``
    (neuquen_basin.py:33-47); ``area = (lat.min(), lat.max(), lon.min(), lon.max())``."""
    path = DEFAULT_THICKNESS_FILE if path is None else path
    lat, lon, thickness = np.loadtxt(path, unpack=True)
    if lat.size != shape[0] * shape[1]:
        raise ValueError("{}: {} rows, expected {} x {}".format(path, lat.size, *shape))
    area = (lat.min(), lat.max(), lon.min(), lon.max())
    return {"lat": lat, "lon": lon, "thickness": thickness, "shape": tuple(shape), "area": area}


def neuquen_basin_model(sediments=None, top=NEUQUEN_TOP):
    """
    ``(basin, top, bottom)``: the tesseroid model of the sediment layer and the highest top /
    lowest bottom of the basin cells, which the reference uses to define its densities
    (neuquen_basin.py:55-72, :93).  `sediments` is the dictionary of
    :func:`load_sediment_thickness`, a path, or ``None`` for the file shipped in ``data/``.
    """
    if sediments is None or isinstance(sediments, (str, bytes, os.PathLike)):
        sediments = load_sediment_thickness(sediments)
    thickness = np.asarray(sediments["thickness"], dtype=float)
    nans = np.isnan(thickness)
    sed_top = np.full(thickness.shape, float(top))
    sed_top[nans] = np.nan
    sed_bottom = sed_top - thickness
    t, b = sed_top.copy(), sed_bottom.copy()
    t[nans] = 0
    b[nans] = 0
    basin = TesseroidModel(sediments["area"], t, b, sediments["shape"])
    return basin, basin.top[~nans].max(), basin.bottom[~nans].min()
