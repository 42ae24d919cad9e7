"""fatiando.mesher.Tesseroid / TesseroidMesh (code/scripts/linear_density.py:6,47)."""
from tesseroid_variable_density_b200.mesher import Tesseroid, TesseroidMesh  # noqa: F401
