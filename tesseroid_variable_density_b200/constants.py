"""
Physical constants used by the forward model.

The reference imports them from Fatiando 0.5 (``fatiando.constants``;
code/tesseroid_density/tesseroid.py:97, _tesseroid.pyx:9,23).  ``G`` is Fatiando's
6.673e-11 (the paper prints 6.674e-11); it is pinned by the bit-exact replay of the
reference's stored results.
"""
#: Gravitational constant (m^3 kg^-1 s^-2)
G = 0.00000000006673
#: m/s^2 -> mGal
SI2MGAL = 100000.0
#: 1/s^2 -> Eotvos
SI2EOTVOS = 1000000000.0
#: Mean Earth radius (m)
MEAN_EARTH_RADIUS = 6378137.0
