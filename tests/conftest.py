"""
pytest configuration.

* ``-m "not gpu"``: oracle vs golden vectors / the unmodified reference, host logic, ABI symbols;
  runs without a GPU.
* ``-m gpu``: the parity tests proper, CUDA path (through the C ABI) vs the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_library_built():
    """The tests load the in-tree libtvd.so; build it (nvcc cross-compiles without a GPU) if a
    fresh checkout has not been through __graft_entry__.build() yet."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "tesseroid_variable_density_b200", "csrc", "libtvd.so")
    if os.path.isfile(lib):
        return
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if os.path.isfile(nvcc):
        subprocess.check_call(["bash", os.path.join(os.path.dirname(lib), "build.sh")],
                              stdout=subprocess.DEVNULL)


def pytest_configure(config):
    _ensure_library_built()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long golden replays (still part of the CPU suite)")


def _has_gpu():
    try:
        from tesseroid_variable_density_b200 import _lib
        return _lib.load().tvd_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    oracle.load()
    return oracle


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (skip if it was not built)."""
    import build_ref
    if not build_ref.available():
        if os.path.isdir("/root/reference/code/tesseroid_density"):
            build_ref.build("/root/reference")
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return build_ref.load()
