"""
Pin of the gradient-tensor integrands (gxx ... gzz) to the REFERENCE.

The reference lists the tensor in its docstring (code/tesseroid_density/tesseroid.py:56-65) but
only implements potential, gx, gy, gz; the tensor kernels live in its un-vendored upstream
(Fatiando 0.5).  What the reference does pin is the gradient vector: for a cell that is not
subdivided its gx, gy, gz are finite sums over the 2x2x2 GLQ nodes of functions of the
computation point (_tesseroid.pyx:309-520), so the derivatives of those sums with respect to the
point ARE the GLQ sums of the tensor integrands -- differentiation commutes with a finite sum over
fixed nodes.  Central differences (Richardson-extrapolated) of the UNMODIFIED reference's
gx/gy/gz, taken along straight lines in an Earth-fixed Cartesian frame and projected on the
point's local North/East/Up axes, therefore give the tensor the reference would compute -- signs,
units (Eotvos), axes (z up) and the off-diagonal terms included, none of which the closed shell
pins (there gxy = gxz = gyz = 0).

The geometry keeps every cell undivided (checked with the oracle's counters) so that the fields
are smooth in the point: the adaptive subdivision only re-applies the same integrand to
sub-cells.
"""
import numpy as np
import pytest

import cases
from tesseroid_variable_density_b200 import density as D
from tesseroid_variable_density_b200.constants import G, MEAN_EARTH_RADIUS, SI2EOTVOS, SI2MGAL
from tesseroid_variable_density_b200.mesher import TesseroidMesh
from tesseroid_variable_density_b200.model import flatten_model
from tesseroid_variable_density_b200.tesseroid import _convert_coords

TENSOR = ("gxx", "gxy", "gxz", "gyy", "gyz", "gzz")


def _frames(lon_deg, lat_deg):
    lam, phi = np.radians(lon_deg), np.radians(lat_deg)
    e_n = np.array([-np.sin(phi) * np.cos(lam), -np.sin(phi) * np.sin(lam), np.cos(phi)])
    e_e = np.array([-np.sin(lam), np.cos(lam), 0.0])
    e_u = np.array([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)])
    return e_n, e_e, e_u


def _to_spherical(xyz):
    r = np.sqrt(np.sum(xyz ** 2))
    lat = np.degrees(np.arcsin(xyz[2] / r))
    lon = np.degrees(np.arctan2(xyz[1], xyz[0]))
    return lon, lat, r - MEAN_EARTH_RADIUS


def _gravity_vector(ref, mesh, xyz):
    """g (m/s^2) in the Earth-fixed frame at the Cartesian point xyz from the reference's
    gx (north), gy (east), gz (DOWN), all in mGal."""
    lon, lat, h = _to_spherical(xyz)
    a = [np.array([v]) for v in (lon, lat, h)]
    gx = ref.gx(*a, mesh, delta=0.1)[0]
    gy = ref.gy(*a, mesh, delta=0.1)[0]
    gz = ref.gz(*a, mesh, delta=0.1)[0]
    e_n, e_e, e_u = _frames(lon, lat)
    return (gx * e_n + gy * e_e - gz * e_u) / SI2MGAL


def _fd_tensor(ref, mesh, lon, lat, height, h):
    e = _frames(lon, lat)
    r = MEAN_EARTH_RADIUS + height
    p0 = r * e[2]
    T = np.zeros((3, 3))
    for j in range(3):
        dg = (_gravity_vector(ref, mesh, p0 + h * e[j]) -
              _gravity_vector(ref, mesh, p0 - h * e[j])) / (2 * h)
        for i in range(3):
            T[i, j] = np.dot(e[i], dg)
    return T * SI2EOTVOS


@pytest.mark.parametrize("dens_name", ["exponential", "sine", "constant"])
def test_tensor_equals_finite_differences_of_reference_gravity(reference, oracle_lib, dens_name):
    mesh = TesseroidMesh((-60.4, -59.6, -30.3, -29.7, 500., -9500.), (1, 6, 8))     # 0.1 deg cells
    dens = {"exponential": D.ExponentialDensity.between(2670, 3300, 500., -9500., 3),
            "sine": D.SineDensity(300., 2 * np.pi * 2 / 10000., 2900.),
            "constant": 2800.}[dens_name]
    mesh.addprop("density", [dens] * mesh.size)
    flat = flatten_model(mesh)
    points = [(-60.0, -30.0, 40e3), (-60.27, -29.82, 45e3), (-59.55, -30.41, 60e3),
              (-60.9, -29.1, 50e3)]
    worst = 0.0
    for lon, lat, height in points:
        cc = _convert_coords(np.array([lon]), np.array([lat]), np.array([height]))
        # nothing is subdivided at these distances, for the vector fields (ratio 2.5) as for
        # the tensor evaluated with the same ratio
        _, st = oracle_lib.forward("gz", flat.bounds, flat.kind, flat.params, 0.1, 2.5, *cc)
        assert st["splits"] == 0
        want = {}
        for f in TENSOR:
            v, st = oracle_lib.forward(f, flat.bounds, flat.kind, flat.params, 0.1, 2.5, *cc)
            assert st["splits"] == 0
            want[f] = v[0] * G * SI2EOTVOS
        t1 = _fd_tensor(reference, mesh, lon, lat, height, 200.0)
        t2 = _fd_tensor(reference, mesh, lon, lat, height, 400.0)
        T = (4 * t1 - t2) / 3                       # Richardson: O(h^4)
        scale = max(abs(v) for v in want.values())
        got = {"gxx": T[0, 0], "gxy": 0.5 * (T[0, 1] + T[1, 0]), "gxz": 0.5 * (T[0, 2] + T[2, 0]),
               "gyy": T[1, 1], "gyz": 0.5 * (T[1, 2] + T[2, 1]), "gzz": T[2, 2]}
        for f in TENSOR:
            err = abs(got[f] - want[f]) / scale
            worst = max(worst, err)
            # tolerance: the finite differences (h = 200/400 m at >= 40 km, extrapolated) are good
            # to ~1e-9; the reference's own gx/gy/gz carry ~1e-12 of rounding, amplified by 1/h
            assert err <= 1e-8, (dens_name, (lon, lat, height), f, got[f], want[f], err)
        # the finite-difference tensor is symmetric and trace-free (Laplace) to the same accuracy
        assert abs(T[0, 1] - T[1, 0]) <= 1e-6 * scale and abs(T[0, 2] - T[2, 0]) <= 1e-6 * scale
        assert abs(T[0, 0] + T[1, 1] + T[2, 2]) <= 1e-6 * scale
        # off-diagonal components are exercised (not a symmetric configuration)
        if (lon, lat) != (-60.0, -30.0):
            assert abs(want["gxy"]) > 1e-3 * scale and abs(want["gyz"]) > 1e-3 * scale
    print("tensor vs finite differences of the reference: worst relative difference %.2e" % worst)
