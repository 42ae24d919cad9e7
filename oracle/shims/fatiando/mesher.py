"""
``Tesseroid`` and ``TesseroidMesh`` with the cell-indexing arithmetic of Fatiando 0.5's
``PrismMesh`` (shape = (nr, nlat, nlon); cell bounds are ``bounds[0] + dims[0]*i`` and that
value ``+ dims[0]``).  Pinned by the golden replay in tests/test_oracle_ref.py.
"""
import numpy as np


class Tesseroid(object):
    def __init__(self, w, e, s, n, top, bottom, props=None):
        self.w, self.e, self.s, self.n = float(w), float(e), float(s), float(n)
        self.top, self.bottom = float(top), float(bottom)
        self.props = {} if props is None else dict(props)

    def addprop(self, prop, value):
        self.props[prop] = value

    def get_bounds(self):
        return [self.w, self.e, self.s, self.n, self.top, self.bottom]


class TesseroidMesh(object):
    celltype = Tesseroid

    def __init__(self, bounds, shape, props=None):
        nz, ny, nx = shape
        x1, x2, y1, y2, z1, z2 = bounds
        self.shape = tuple(int(i) for i in shape)
        self.size = int(nx*ny*nz)
        self.dims = ((x2 - x1)/nx, (y2 - y1)/ny, (z2 - z1)/nz)
        self.bounds = bounds
        self.props = {} if props is None else props
        self.i = 0
        self.mask = []

    def __len__(self):
        return self.size

    def __getitem__(self, index):
        if index >= self.size or index < -self.size:
            raise IndexError('mesh index out of range')
        if index < 0:
            index = self.size + index
        if index in self.mask:
            return None
        nz, ny, nx = self.shape
        k = index//(nx*ny)
        j = (index - k*(nx*ny))//nx
        i = index - k*(nx*ny) - j*nx
        x1 = self.bounds[0] + self.dims[0]*i
        x2 = x1 + self.dims[0]
        y1 = self.bounds[2] + self.dims[1]*j
        y2 = y1 + self.dims[1]
        z1 = self.bounds[4] + self.dims[2]*k
        z2 = z1 + self.dims[2]
        props = dict([p, self.props[p][index]] for p in self.props)
        return self.celltype(x1, x2, y1, y2, z1, z2, props=props)

    def __iter__(self):
        self.i = 0
        return self

    def __next__(self):
        if self.i >= self.size:
            raise StopIteration
        cell = self.__getitem__(self.i)
        self.i += 1
        return cell

    next = __next__

    def addprop(self, prop, values):
        self.props[prop] = values
