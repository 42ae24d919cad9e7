"""Stand-in for the ``future`` package (tesseroid.py:91 imports ``future.builtins.range``)."""
