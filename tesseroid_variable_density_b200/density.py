"""
Density-in-depth models.

The reference accepts any Python callable ``density(height)`` and calls it 8 times per
integrated cell from Cython (code/tesseroid_density/_tesseroid.pyx:274,354,434,515).  A Python
callable cannot run on a GPU, so the drop-in takes *density objects*: each carries the
parameters the CUDA kernels evaluate and is ALSO a callable with the reference scripts' exact
expression order, so the same object can be handed to the reference implementation.

Plain numbers stay supported (constant density, _tesseroid.pyx:218-224).  Arbitrary callables
raise ``TypeError``: there is no CPU fallback.
"""
from __future__ import annotations

import numbers

import numpy as np

from .constants import MEAN_EARTH_RADIUS

#: density kinds shared with include/tvd.h (enum tvd_density_kind)
CONST, LINEAR, EXP, SINE = 0, 1, 2, 3
NPAR = 5


class Density(object):
    """Base class: ``kind`` (int), ``params`` (5 floats) and ``__call__(height)``."""
    kind = -1

    @property
    def params(self):
        raise NotImplementedError

    def __call__(self, height):
        raise NotImplementedError

    def __repr__(self):
        return "{}({})".format(type(self).__name__,
                               ", ".join(repr(float(p)) for p in self.params))


class LinearDensity(Density):
    r"""
    :math:`\rho(h) = slope\,(h + offset) + constant`.

    With ``offset = MEAN_EARTH_RADIUS`` this is ``density_linear`` of
    code/scripts/linear_density.py:73-75 (``slope * r + constant_term``); with
    ``offset = -bottom`` it is ``linear_density`` of code/scripts/neuquen_basin.py:119-121
    (``slope * (h - bottom) + density_bottom``).
    """
    kind = LINEAR

    def __init__(self, slope, offset=0.0, constant=0.0):
        self.slope, self.offset, self.constant = float(slope), float(offset), float(constant)

    @property
    def params(self):
        return (self.slope, self.offset, self.constant, 0.0, 0.0)

    def __call__(self, height):
        r = height + self.offset
        return self.slope * r + self.constant

    @classmethod
    def from_radius(cls, slope, constant_term):
        """``slope*(height + MEAN_EARTH_RADIUS) + constant_term`` (linear_density.py:73-75)."""
        return cls(slope, MEAN_EARTH_RADIUS, constant_term)

    @classmethod
    def between(cls, density_top, density_bottom, top, bottom):
        """Linear from ``density_bottom`` at ``bottom`` to ``density_top`` at ``top``
        (neuquen_basin.py:119-121)."""
        slope = (density_top - density_bottom) / (top - bottom)
        return cls(slope, -bottom, density_bottom)


class ExponentialDensity(Density):
    r"""
    :math:`\rho(h) = A\,\exp(rate\,(h - h_0)/scale) + C`, evaluated as
    ``A * np.exp(rate * (h - h0) / scale) + C``: the expression of
    code/scripts/exponential_density.py:97-98 and neuquen_basin.py:142-147 with
    ``rate = -b_factor``, ``h0 = bottom``, ``scale = thickness``; the notebook's
    ``a*np.exp(-height/b) + c`` (cell 6) is ``rate=-1, h0=0, scale=b``.
    """
    kind = EXP

    def __init__(self, amplitude, rate, h0, scale, constant):
        self.amplitude, self.rate, self.h0 = float(amplitude), float(rate), float(h0)
        self.scale, self.constant = float(scale), float(constant)

    @property
    def params(self):
        return (self.amplitude, self.rate, self.h0, self.scale, self.constant)

    def __call__(self, height):
        return self.amplitude * np.exp(self.rate * (height - self.h0) / self.scale) + self.constant

    @classmethod
    def between(cls, density_top, density_bottom, top, bottom, b_factor):
        """The scripts' parametrisation: ``A = (rho_bottom - rho_top)/(1 - exp(-b))``,
        ``C = rho_bottom - A``, ``rho = A*exp(-b*(h - bottom)/thickness) + C``
        (exponential_density.py:93-98)."""
        thickness = top - bottom
        A = (density_bottom - density_top) / (1 - np.exp(-b_factor))
        C = density_bottom - A
        return cls(A, -b_factor, bottom, thickness, C)


class SineDensity(Density):
    r"""
    :math:`\rho(h) = A\,\sin(k\,h) + C`, evaluated as ``A * np.sin(k * h) + C``
    (code/scripts/sine_density.py:93-94 with ``A = C = max_density/2``).
    """
    kind = SINE

    def __init__(self, amplitude, k_factor, constant):
        self.amplitude, self.k_factor = float(amplitude), float(k_factor)
        self.constant = float(constant)

    @property
    def params(self):
        return (self.amplitude, self.k_factor, self.constant, 0.0, 0.0)

    def __call__(self, height):
        return self.amplitude * np.sin(self.k_factor * height) + self.constant


def as_kind_params(density):
    """
    Map a density specification to ``(kind, params)`` for the C ABI.

    A number is a constant density; a :class:`Density` carries its own parameters; anything
    else (e.g. an arbitrary Python callable) cannot be evaluated on the GPU.
    """
    if isinstance(density, Density):
        return density.kind, density.params
    if isinstance(density, (numbers.Real, np.floating, np.integer)) and not isinstance(density, bool):
        return CONST, (float(density), 0.0, 0.0, 0.0, 0.0)
    if callable(density):
        raise TypeError(
            "arbitrary Python callables cannot be evaluated on the GPU; use a density object "
            "(LinearDensity, ExponentialDensity, SineDensity) or a number. There is no CPU "
            "fallback.")
    raise TypeError("unsupported density {!r}".format(density))
