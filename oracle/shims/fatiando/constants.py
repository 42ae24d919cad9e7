"""
The four ``fatiando.constants`` values the reference path reads
(/root/reference/code/tesseroid_density/tesseroid.py:97, _tesseroid.pyx:9,23).
Values are those of Fatiando 0.5; G is pinned by the bit-exact replay of the reference's
stored golden results.
"""
#: Gravitational constant (m^3 kg^-1 s^-2) as shipped by Fatiando 0.5
G = 0.00000000006673
#: m/s^2 -> mGal
SI2MGAL = 100000.0
#: 1/s^2 -> Eotvos
SI2EOTVOS = 1000000000.0
#: Mean Earth radius in metres
MEAN_EARTH_RADIUS = 6378137.0
