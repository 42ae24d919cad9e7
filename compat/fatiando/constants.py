"""fatiando.constants (code/tesseroid_density/tesseroid.py:97)."""
from tesseroid_variable_density_b200.constants import *  # noqa: F401,F403
from tesseroid_variable_density_b200.constants import (G, MEAN_EARTH_RADIUS, SI2EOTVOS,  # noqa: F401
                                                       SI2MGAL)
