"""``from tesseroid_density import tesseroid`` -- the reference's import path
(code/scripts/linear_density.py:11), served by the B200 library."""
from . import tesseroid  # noqa: F401
