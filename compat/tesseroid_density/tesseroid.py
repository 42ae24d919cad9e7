"""
The call surface of ``tesseroid_density.tesseroid`` (code/tesseroid_density/tesseroid.py) over
tesseroid_variable_density_b200, for callers that were written against the reference.

The one thing the reference allows that a GPU cannot do is calling back into Python for the
density (_tesseroid.pyx:274).  Here callables are recognised once per distinct function with
``density.fit_density`` (constant / linear / exponential / sinusoidal in height: every density
the reference's scripts define) and replaced by density objects; anything else raises
``TypeError``.
"""
from __future__ import annotations

import numpy as np

from tesseroid_variable_density_b200 import tesseroid as _t
from tesseroid_variable_density_b200.density import Density, fit_density
from tesseroid_variable_density_b200.model import DensityTable, FlatModel
from tesseroid_variable_density_b200.tesseroid import (DELTA, RATIO_G, RATIO_GG, RATIO_V,  # noqa: F401
                                                       STACK_SIZE, PreparedModel,
                                                       _check_input, _convert_coords,
                                                       _split_arrays)


def _is_plain_callable(d):
    return callable(d) and not isinstance(d, Density)


def _height_range(model):
    if hasattr(model, "flat_bounds"):
        b, keep = model.flat_bounds()
        b = b[keep]
    else:
        b = np.array([c.get_bounds() for c in model if c is not None], dtype=float).reshape(-1, 6)
    return float(np.max(b[:, 4])), float(np.min(b[:, 5]))


class _Adapted(object):
    """A view of a model whose callable densities are replaced by fitted density objects."""

    def __init__(self, model, props):
        self._model = model
        self.props = props

    def flat_bounds(self):
        return self._model.flat_bounds()


def _adapt(model, dens):
    """(model, dens) with plain callables replaced by density objects."""
    if isinstance(model, (PreparedModel, FlatModel)):
        return model, dens
    fitted = {}
    rng = []

    def conv(d):
        if not _is_plain_callable(d):
            return d
        key = id(d)
        if key not in fitted:
            if not rng:
                rng.extend(_height_range(model))
            fitted[key] = fit_density(d, rng[0], rng[1])
        return fitted[key]

    if dens is not None:
        return model, conv(dens)
    if hasattr(model, "flat_bounds") and hasattr(model, "props"):
        seq = model.props.get("density")
        if seq is None or isinstance(seq, (DensityTable, np.ndarray)):
            return model, dens
        seq = list(seq)
        if not any(_is_plain_callable(d) for d in {id(d): d for d in seq}.values()):
            return model, dens
        props = dict(model.props)
        props["density"] = [conv(d) for d in seq]
        return _Adapted(model, props), dens
    cells = list(model)
    if not any(c is not None and _is_plain_callable(c.props.get("density")) for c in cells):
        return cells, dens
    out = []
    for c in cells:
        if c is not None and _is_plain_callable(c.props.get("density")):
            from tesseroid_variable_density_b200.mesher import Tesseroid
            p = dict(c.props)
            if not rng:
                b = np.array([x.get_bounds() for x in cells if x is not None], dtype=float)
                rng.extend((float(b[:, 4].max()), float(b[:, 5].min())))
            p["density"] = conv(p["density"])
            c = Tesseroid(*c.get_bounds(), props=p)
        out.append(c)
    return out, dens


def _wrap(name):
    func = getattr(_t, name)

    def wrapped(lon, lat, height, model, dens=None, **kwargs):
        model, dens = _adapt(model, dens)
        return func(lon, lat, height, model, dens=dens, **kwargs)
    wrapped.__name__ = name
    wrapped.__doc__ = func.__doc__
    return wrapped


potential, gx, gy, gz = _wrap("potential"), _wrap("gx"), _wrap("gy"), _wrap("gz")
gxx, gxy, gxz = _wrap("gxx"), _wrap("gxy"), _wrap("gxz")
gyy, gyz, gzz = _wrap("gyy"), _wrap("gyz"), _wrap("gzz")


def _density_based_discretization(bounds, density, delta):
    """Same contract as the reference (tesseroid.py:237-283); callables are fitted first."""
    if _is_plain_callable(density):
        density = fit_density(density, bounds[4], bounds[5])
    return _t._density_based_discretization(bounds, density, delta)
