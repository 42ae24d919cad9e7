"""The names of Fatiando a Terra 0.5 that pinga-lab/tesseroid-variable-density imports
(code/scripts/linear_density.py:5-7), provided by tesseroid_variable_density_b200."""
from . import constants, gridder, mesher  # noqa: F401
