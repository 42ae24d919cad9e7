"""
tesseroid_variable_density_b200 -- the forward-modelling hot path of
pinga-lab/tesseroid-variable-density (tesseroids with density varying in depth), built from
scratch for NVIDIA B200 (sm_100a): hand-written FP64 CUDA kernels behind a C ABI
(``include/tvd.h``), driven from Python through ctypes with the reference's call surface.

    from tesseroid_variable_density_b200 import tesseroid
    gz = tesseroid.gz(lon, lat, height, model, ratio=2.5, delta=0.1)
"""
from . import constants, datasets, density, gridder, mesher, model  # noqa: F401
from .density import LinearDensity, ExponentialDensity, SineDensity  # noqa: F401
from .mesher import Tesseroid, TesseroidMesh, TesseroidModel        # noqa: F401
from . import tesseroid                                            # noqa: F401

__all__ = ["tesseroid", "density", "mesher", "gridder", "constants", "model", "datasets",
           "LinearDensity", "ExponentialDensity", "SineDensity",
           "Tesseroid", "TesseroidMesh", "TesseroidModel"]
