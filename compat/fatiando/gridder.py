"""fatiando.gridder.regular (code/scripts/linear_density.py:54-57)."""
from tesseroid_variable_density_b200.gridder import regular  # noqa: F401
