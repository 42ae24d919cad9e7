"""
Flattening of a tesseroid model into the struct-of-arrays form the C ABI takes.

Restates the per-tesseroid checks of the reference's model loop
(``_check_tesseroid``, code/tesseroid_density/tesseroid.py:142-161 and :217-221): ``None``
cells and cells without a ``'density'`` property are skipped (unless ``dens`` is given, which
then overrides every cell's density), and ``w <= e and s <= n and top >= bottom`` is asserted.
"""
from __future__ import annotations

import numbers

import numpy as np

from . import density as _density


class DensityTable(object):
    """
    Columnar sequence of per-cell densities of one kind: ``kind`` (int or int array of length
    n) and ``params`` ((n, 5) array).  Behaves like a list of density objects (indexing returns
    a :class:`~.density.Density`), but flattens without creating n Python objects.
    """

    def __init__(self, kind, params):
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        assert self.params.ndim == 2 and self.params.shape[1] == _density.NPAR
        n = self.params.shape[0]
        self.kind = np.ascontiguousarray(np.broadcast_to(np.asarray(kind, dtype=np.int32), (n,)))

    def __len__(self):
        return self.params.shape[0]

    def __getitem__(self, index):
        k, p = int(self.kind[index]), self.params[index]
        if k == _density.CONST:
            return float(p[0])
        if k == _density.LINEAR:
            return _density.LinearDensity(p[0], p[1], p[2])
        if k == _density.EXP:
            return _density.ExponentialDensity(*p)
        return _density.SineDensity(p[0], p[1], p[2])


class FlatModel(object):
    """``bounds`` (n, 6) float64, ``kind`` (n,) int32, ``params`` (n, 5) float64."""

    def __init__(self, bounds, kind, params):
        self.bounds = np.ascontiguousarray(bounds, dtype=np.float64).reshape(-1, 6)
        self.kind = np.ascontiguousarray(kind, dtype=np.int32).reshape(-1)
        self.params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, _density.NPAR)
        assert self.bounds.shape[0] == self.kind.shape[0] == self.params.shape[0]

    @property
    def size(self):
        return self.bounds.shape[0]


def _check_bounds(bounds):
    ok = (bounds[:, 0] <= bounds[:, 1]) & (bounds[:, 2] <= bounds[:, 3]) & \
        (bounds[:, 4] >= bounds[:, 5])
    if not np.all(ok):
        bad = int(np.argmin(ok))
        # same assertion (and message) as tesseroid.py:155-156
        raise AssertionError("Invalid tesseroid dimensions {}".format(list(bounds[bad])))


def _is_number(d):
    return isinstance(d, (numbers.Real, np.floating, np.integer)) and not isinstance(d, bool)


def _densities_to_arrays(dens_seq, n):
    """kind/params arrays (and a keep mask) for a per-cell density sequence: cell ``i`` has
    density ``dens_seq[i]`` (the reference indexes ``props['density'][index]``,
    fatiando.mesher); ``None`` entries are skipped like cells without a density
    (tesseroid.py:149-152)."""
    keep = np.ones(n, dtype=bool)
    if isinstance(dens_seq, DensityTable):
        assert len(dens_seq) == n
        return dens_seq.kind, dens_seq.params, keep
    if n == 0:
        return np.zeros(0, dtype=np.int32), np.zeros((0, _density.NPAR)), keep
    assert len(dens_seq) >= n, "fewer densities ({}) than cells ({})".format(len(dens_seq), n)
    # a numeric array: constant densities, vectorised
    if isinstance(dens_seq, np.ndarray) and dens_seq.dtype != object:
        if not np.issubdtype(dens_seq.dtype, np.number):
            raise TypeError("unsupported density array of dtype {}".format(dens_seq.dtype))
        params = np.zeros((n, _density.NPAR))
        params[:, 0] = np.asarray(dens_seq, dtype=np.float64).reshape(len(dens_seq), -1)[:n, 0]
        return np.zeros(n, dtype=np.int32), params, keep
    # anything else is held as a list for the rest of this function: object identities are only
    # meaningful while the objects are alive (iterating an array creates temporaries)
    seq = dens_seq if isinstance(dens_seq, (list, tuple)) else list(dens_seq)
    if len(seq) > n:
        seq = seq[:n]
    first = seq[0]
    if first is not None and seq.count(first) == n:
        # one density shared by every cell (the scripts' ``[density] * mesh.size``)
        k, p = _density.as_kind_params(first)
        return (np.full(n, k, dtype=np.int32), np.tile(np.asarray(p, dtype=np.float64), (n, 1)),
                keep)
    if _is_number(first):
        try:
            arr = np.array(seq, dtype=np.float64)
        except (TypeError, ValueError):
            arr = None
        if arr is not None and arr.shape == (n,) and not any(d is None for d in seq):
            params = np.zeros((n, _density.NPAR))
            params[:, 0] = arr
            return np.zeros(n, dtype=np.int32), params, keep
    # objects: convert every DISTINCT object once (a mesh of a million cells usually shares a
    # handful of density objects) and scatter with NumPy
    ids = np.fromiter(map(id, seq), dtype=np.int64, count=n)
    uniq, first_at, inverse = np.unique(ids, return_index=True, return_inverse=True)
    tab_kind = np.zeros(uniq.size, dtype=np.int32)
    tab_params = np.zeros((uniq.size, _density.NPAR))
    tab_keep = np.ones(uniq.size, dtype=bool)
    for u, i in enumerate(first_at):
        d = seq[int(i)]
        if d is None:
            tab_keep[u] = False
            continue
        k, p = _density.as_kind_params(d)
        tab_kind[u] = k
        tab_params[u] = p
    inverse = inverse.reshape(-1)
    return tab_kind[inverse], tab_params[inverse], keep & tab_keep[inverse]


def flatten_model(model, dens=None):
    """
    Flatten ``model`` (any iterable following the reference's model protocol) into a
    :class:`FlatModel`.  Containers offering ``flat_bounds()`` and a ``props`` dictionary
    (:class:`~.mesher.TesseroidMesh`, :class:`~.mesher.TesseroidModel`) take a vectorised path.
    """
    if isinstance(model, FlatModel):
        if dens is None:
            return model
        k, p = _density.as_kind_params(dens)
        n = model.size
        return FlatModel(model.bounds, np.full(n, k, dtype=np.int32), np.tile(p, (n, 1)))
    if hasattr(model, "flat_bounds") and hasattr(model, "props"):
        bounds, keep = model.flat_bounds()
        n = bounds.shape[0]
        if dens is not None:
            k, p = _density.as_kind_params(dens)
            kind = np.full(n, k, dtype=np.int32)
            params = np.tile(np.asarray(p, dtype=float), (n, 1))
        elif "density" in model.props:
            kind, params, keep2 = _densities_to_arrays(model.props["density"], n)
            keep = keep & keep2
        else:
            keep = np.zeros(n, dtype=bool)
            kind = np.zeros(n, dtype=np.int32)
            params = np.zeros((n, _density.NPAR))
        if not keep.all():
            bounds, kind, params = bounds[keep], kind[keep], params[keep]
        _check_bounds(bounds)
        return FlatModel(bounds, kind, params)
    # generic iterable: the reference's loop (tesseroid.py:217-221)
    bounds, kinds, params = [], [], []
    cache = {}
    override = _density.as_kind_params(dens) if dens is not None else None
    for tesseroid in model:
        if tesseroid is None:
            continue
        if "density" not in tesseroid.props and dens is None:
            continue
        b = tesseroid.get_bounds()
        w, e, s, n, top, bottom = b
        assert w <= e and s <= n and top >= bottom, \
            "Invalid tesseroid dimensions {}".format(tesseroid.get_bounds())
        if override is not None:
            kp = override
        else:
            d = tesseroid.props["density"]
            # the cache holds the object itself: an id is only unique while its object is alive
            # (a generator may build a fresh density per tesseroid)
            hit = cache.get(id(d))
            if hit is not None:
                kp = hit[1]
            else:
                kp = _density.as_kind_params(d)
                if isinstance(d, _density.Density):
                    cache[id(d)] = (d, kp)
        bounds.append(b)
        kinds.append(kp[0])
        params.append(kp[1])
    return FlatModel(np.array(bounds, dtype=np.float64).reshape(-1, 6),
                     np.array(kinds, dtype=np.int32),
                     np.array(params, dtype=np.float64).reshape(-1, _density.NPAR))
