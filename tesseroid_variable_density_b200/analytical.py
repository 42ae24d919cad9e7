"""
Closed-form gravitational fields of a spherical shell whose density depends on the radius
only, at a point outside it.  These are the known answers the reference validates against
(code/scripts/linear_density.py:14-31, exponential_density.py:14-48, sine_density.py:14-42;
derivations in manuscript/manuscript.tex:1374-1494):

    V = (4 pi G / r) * integral rho(r') r'^2 dr',   gz = V/r,
    gxx = gyy = -V/r^2,   gzz = 2 V/r^2,   all other components vanish

in SI, mGal and Eotvos respectively (gz with z pointing down).
"""
from __future__ import annotations

import numpy as np

from .constants import G, MEAN_EARTH_RADIUS, SI2MGAL, SI2EOTVOS


def _fields_from_potential(potential, r):
    return {'potential': potential,
            'gx': 0,
            'gy': 0,
            'gz': SI2MGAL*(potential/r),
            'gxx': SI2EOTVOS*(-potential/r**2),
            'gxy': 0,
            'gxz': 0,
            'gyy': SI2EOTVOS*(-potential/r**2),
            'gyz': 0,
            'gzz': SI2EOTVOS*(2*potential/r**2)}


def shell_linear_density(height, top, bottom, slope, constant_term):
    """Shell with rho(r') = slope*r' + constant_term (linear_density.py:14-31)."""
    r = height + MEAN_EARTH_RADIUS
    r1 = bottom + MEAN_EARTH_RADIUS
    r2 = top + MEAN_EARTH_RADIUS
    constant = np.pi * G * slope * (r2**4 - r1**4) + \
        4/3. * np.pi * G * constant_term * (r2**3 - r1**3)
    return _fields_from_potential(constant/r, r)


def shell_exponential_density(height, top, bottom, amplitude, b_factor, constant_term):
    """Shell with rho(r') = A exp(-b (r' - R1)/T) + C (exponential_density.py:14-48)."""
    r = height + MEAN_EARTH_RADIUS
    r1 = bottom + MEAN_EARTH_RADIUS
    r2 = top + MEAN_EARTH_RADIUS
    thickness = float(top - bottom)
    k = b_factor / thickness
    potential = 4 * np.pi * G * amplitude / k**3 / r * \
        (
         ((r1 * k)**2 + 2*r1*k + 2) -
         ((r2 * k)**2 + 2*r2*k + 2) * np.exp(-k * thickness)
        ) + \
        4 / 3. * np.pi * G * constant_term * (r2**3 - r1**3) / r
    return _fields_from_potential(potential, r)


def shell_sine_density(height, top, bottom, A, k, constant_density):
    """Shell with rho(r') = A sin(k (r' - R)) + C (sine_density.py:14-42)."""
    r = height + MEAN_EARTH_RADIUS
    r1 = bottom + MEAN_EARTH_RADIUS
    r2 = top + MEAN_EARTH_RADIUS
    potential = 4 * np.pi * G * A / r / k**3 * (
        (2 - k**2 * r2**2) * np.cos(k * (r2 - MEAN_EARTH_RADIUS)) +
        2 * k * r2 * np.sin(k * (r2 - MEAN_EARTH_RADIUS)) -
        (2 - k**2 * r1**2) * np.cos(k * (r1 - MEAN_EARTH_RADIUS)) -
        2 * k * r1 * np.sin(k * (r1 - MEAN_EARTH_RADIUS))
        ) + \
        4 / 3. * np.pi * G * constant_density * (r2**3 - r1**3) / r
    return _fields_from_potential(potential, r)
