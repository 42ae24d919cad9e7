"""
Tesseroid containers that satisfy the reference's model protocol: any iterable whose elements
are ``None`` or have ``.props`` (dict with key ``'density'``) and
``.get_bounds() -> [w, e, s, n, top, bottom]`` (code/tesseroid_density/tesseroid.py:149-160).

* :class:`Tesseroid`, :class:`TesseroidMesh` -- the two Fatiando 0.5 classes the reference
  scripts use (code/scripts/linear_density.py:6,47; notebook cells 4, 12).  Cell bounds use the
  same arithmetic as Fatiando's ``PrismMesh`` (``bounds[0] + dims[0]*i`` and that ``+ dims[0]``).
* :class:`TesseroidModel` -- the layer-with-relief model of the Neuquen application
  (code/scripts/tesseroid_model.py:22-161), including its ``s = lats[i] - dlon/2`` quirk (:146).

Each container also offers ``flat_bounds()``, a vectorised way to the ``(n, 6)`` bounds array
that :mod:`.model` uses instead of iterating a million Python objects.
"""
from __future__ import annotations

import numpy as np


class Tesseroid(object):
    """A tesseroid (spherical prism): west, east, south, north in degrees, top/bottom heights
    in metres, and a ``props`` dictionary."""

    def __init__(self, w, e, s, n, top, bottom, props=None):
        self.w, self.e, self.s, self.n = float(w), float(e), float(s), float(n)
        self.top, self.bottom = float(top), float(bottom)
        self.props = {} if props is None else dict(props)

    def addprop(self, prop, value):
        self.props[prop] = value

    def get_bounds(self):
        return [self.w, self.e, self.s, self.n, self.top, self.bottom]

    def __repr__(self):
        return "Tesseroid(w={}, e={}, s={}, n={}, top={}, bottom={}, props={})".format(
            self.w, self.e, self.s, self.n, self.top, self.bottom, self.props)


class TesseroidMesh(object):
    """
    Regular 3-D mesh of tesseroids.  ``bounds = (w, e, s, n, top, bottom)``,
    ``shape = (nr, nlat, nlon)``; cell ``index`` has ``k = index // (nlon*nlat)`` (radial),
    ``j`` (latitude) and ``i`` (longitude, fastest).
    """
    celltype = Tesseroid

    def __init__(self, bounds, shape, props=None):
        nr, nlat, nlon = (int(v) for v in shape)
        w, e, s, n, top, bottom = bounds
        self.shape = (nr, nlat, nlon)
        self.size = nr * nlat * nlon
        self.dims = ((e - w) / nlon, (n - s) / nlat, (bottom - top) / nr)
        self.bounds = bounds
        self.props = {} if props is None else props
        self.mask = []
        self._i = 0

    def __len__(self):
        return self.size

    def addprop(self, prop, values):
        self.props[prop] = values

    def _cell_bounds(self, index):
        nr, nlat, nlon = self.shape
        k = index // (nlon * nlat)
        j = (index - k * (nlon * nlat)) // nlon
        i = index - k * (nlon * nlat) - j * nlon
        w = self.bounds[0] + self.dims[0] * i
        e = w + self.dims[0]
        s = self.bounds[2] + self.dims[1] * j
        n = s + self.dims[1]
        top = self.bounds[4] + self.dims[2] * k
        bottom = top + self.dims[2]
        return w, e, s, n, top, bottom

    def __getitem__(self, index):
        if index >= self.size or index < -self.size:
            raise IndexError("mesh index out of range")
        if index < 0:
            index = self.size + index
        if index in self.mask:
            return None
        props = dict([p, self.props[p][index]] for p in self.props)
        return self.celltype(*self._cell_bounds(index), props=props)

    def __iter__(self):
        self._i = 0
        return self

    def __next__(self):
        if self._i >= self.size:
            raise StopIteration
        cell = self.__getitem__(self._i)
        self._i += 1
        return cell

    next = __next__

    def flat_bounds(self):
        """``(size, 6)`` float64 array of every cell's bounds (same arithmetic as
        ``__getitem__``) and a boolean mask of the cells that are not masked out."""
        # _cell_bounds' expressions once per column / row / layer, broadcast over the mesh
        nr, nlat, nlon = self.shape
        w = float(self.bounds[0]) + float(self.dims[0]) * np.arange(nlon)
        s = float(self.bounds[2]) + float(self.dims[1]) * np.arange(nlat)
        top = float(self.bounds[4]) + float(self.dims[2]) * np.arange(nr)
        out = np.empty((self.size, 6))
        cells = out.reshape(nr, nlat, nlon, 6)
        cells[..., 0] = w
        cells[..., 1] = w + float(self.dims[0])
        cells[..., 2] = s[:, np.newaxis]
        cells[..., 3] = (s + float(self.dims[1]))[:, np.newaxis]
        cells[..., 4] = top[:, np.newaxis, np.newaxis]
        cells[..., 5] = (top + float(self.dims[2]))[:, np.newaxis, np.newaxis]
        keep = np.ones(self.size, dtype=bool)
        if len(self.mask):
            keep[np.asarray(self.mask, dtype=int)] = False
        return out, keep


class TesseroidModel(object):
    """
    One layer of ``nlat x nlon`` tesseroids with per-cell ``top`` and ``bottom`` (1-D arrays,
    latitude-major).  ``area = (s, n, w, e)``; cell centres are the grid nodes, as in the
    reference's ``TesseroidModel`` (code/scripts/tesseroid_model.py:47-68,139-154).
    """

    def __init__(self, area, top, bottom, shape, props=None):
        top = np.asarray(top, dtype=float)
        bottom = np.asarray(bottom, dtype=float)
        assert shape[0] * shape[1] == top.size
        assert top.size == bottom.size
        assert len(area) == 4
        assert area[0] < area[1] and area[2] < area[3]
        self.area = area
        self.shape = tuple(shape)
        s, n, w, e = area
        nlat, nlon = shape
        self.lons = np.linspace(w, e, nlon, endpoint=True)
        self.lats = np.linspace(s, n, nlat, endpoint=True)
        self.spacing = self.lats[1] - self.lats[0], self.lons[1] - self.lons[0]
        # cells whose top is below their bottom are flipped (tesseroid_model.py:102-110)
        below = top <= bottom
        self._top = np.where(below, bottom, top)
        self._bottom = np.where(below, top, bottom)
        self.props = {} if props is None else props
        self._i = 0

    @property
    def top(self):
        return self._top

    @property
    def bottom(self):
        return self._bottom

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    def __len__(self):
        return self.size

    def addprop(self, prop, values):
        self.props[prop] = values

    def __iter__(self):
        self._i = 0
        return self

    def __next__(self):
        if self._i >= self.size:
            raise StopIteration
        cell = self.__getitem__(self._i)
        self._i += 1
        return cell

    next = __next__

    def __getitem__(self, index):
        nlat, nlon = self.shape
        dlat, dlon = self.spacing
        i = index // nlon
        j = index - i * nlon
        w = self.lons[j] - dlon / 2.
        e = w + dlon
        s = self.lats[i] - dlon / 2.      # sic: dlon, as in tesseroid_model.py:146
        n = s + dlat
        props = dict([p, self.props[p][index]] for p in self.props)
        return Tesseroid(w, e, s, n, self.top[index], self.bottom[index], props)

    def flat_bounds(self):
        # __getitem__'s expressions, evaluated once per column / row and broadcast over the layer
        nlat, nlon = self.shape
        dlat, dlon = self.spacing
        w = self.lons - dlon / 2.
        s = self.lats - dlon / 2.        # sic: dlon, as in __getitem__
        out = np.empty((self.size, 6))
        cells = out.reshape(nlat, nlon, 6)
        cells[:, :, 0] = w[np.newaxis, :]
        cells[:, :, 1] = (w + dlon)[np.newaxis, :]
        cells[:, :, 2] = s[:, np.newaxis]
        cells[:, :, 3] = (s + dlat)[:, np.newaxis]
        out[:, 4] = self.top
        out[:, 5] = self.bottom
        return out, np.ones(self.size, dtype=bool)
