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


def fit_density(func, top, bottom, rtol=1e-9):
    """
    Recognise a Python callable ``func(height)`` as one of the parametric density models on
    ``[bottom, top]`` and return the corresponding number or :class:`Density` object.

    The reference takes arbitrary callables (code/tesseroid_density/_tesseroid.pyx:274); the
    GPU kernels need parameters.  This helper samples the callable, fits a constant, a linear,
    an exponential (``A*exp(k*h) + C``) or a sinusoidal (``A*sin(k*h) + C``) model -- the four
    families of the reference's scripts -- and accepts the first one that reproduces the
    callable to `rtol` (relative to its largest magnitude) on 201 heights.  Raises ``TypeError``
    otherwise.  The fitted parameters agree with the callable's own to ~1e-12, so results differ
    from a run with the exact density object at that level; where bit-level agreement with the
    reference matters, pass the density object itself.
    """
    top, bottom = float(top), float(bottom)
    if not top > bottom:
        raise ValueError("fit_density needs top > bottom")
    h = np.linspace(bottom, top, 201)
    try:
        f = np.asarray(func(h), dtype=np.float64)
        if f.shape != h.shape:
            raise ValueError
    except Exception:
        f = np.array([float(func(float(v))) for v in h])
    scale = float(np.max(np.abs(f)))
    if not np.all(np.isfinite(f)):
        raise TypeError("density callable returns non-finite values on [bottom, top]")

    def ok(model):
        return float(np.max(np.abs(model(h) - f))) <= rtol * max(scale, 1e-300)

    if np.all(f == f[0]):
        return float(f[0])
    lin = LinearDensity((f[-1] - f[0]) / (top - bottom), -bottom, f[0])
    if ok(lin):
        return lin
    # A*exp(k*h) + C from three equally spaced samples: (f2 - f1)/(f1 - f0) = exp(k*dh)
    i0, i1, i2 = 0, 100, 200
    d1, d2 = f[i1] - f[i0], f[i2] - f[i1]
    if d1 != 0 and d2 / d1 > 0 and d2 != d1:
        k = np.log(d2 / d1) / (h[i1] - h[i0])
        amp = d1 / np.expm1(k * (h[i1] - h[i0]))         # A*exp(k*bottom)
        cand = ExponentialDensity(amp, k * (top - bottom), bottom, top - bottom, f[i0] - amp)
        if ok(cand):
            return cand
        try:
            from scipy.optimize import least_squares
            sol = least_squares(lambda p: p[0] * np.exp(p[1] * (h - bottom) / (top - bottom)) + p[2] - f,
                                [amp, k * (top - bottom), f[i0] - amp], xtol=1e-15, ftol=1e-15,
                                gtol=1e-15)
            cand = ExponentialDensity(sol.x[0], sol.x[1], bottom, top - bottom, sol.x[2])
            if ok(cand):
                return cand
        except Exception:
            pass
    # A*sin(k*h) + C: dominant frequency from the spectrum, then least squares
    try:
        from scipy.optimize import least_squares
        g = f - f.mean()
        spec = np.abs(np.fft.rfft(g * np.hanning(g.size), 8 * g.size))
        freq = np.fft.rfftfreq(8 * g.size, d=h[1] - h[0])
        k0 = 2 * np.pi * freq[int(np.argmax(spec[1:])) + 1]
        best = None
        for sign in (1.0, -1.0):
            sol = least_squares(lambda p: p[0] * np.sin(p[1] * h) + p[2] - f,
                                [sign * 0.5 * (f.max() - f.min()), k0, f.mean()], xtol=1e-15,
                                ftol=1e-15, gtol=1e-15, x_scale=[scale, k0, scale])
            cand = SineDensity(sol.x[0], sol.x[1], sol.x[2])
            if ok(cand) and (best is None):
                best = cand
        if best is not None:
            return best
    except Exception:
        pass
    raise TypeError(
        "the density callable is not a constant, linear, exponential (A*exp(k*h) + C) or "
        "sinusoidal (A*sin(k*h) + C) function of height on [{}, {}]; arbitrary callables "
        "cannot be evaluated on the GPU and there is no CPU fallback".format(bottom, top))
