import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    import refharness
    return refharness.load()


@pytest.fixture(scope="session")
def rb():
    """The product library, bound to cuda:0 (GPU tests only)."""
    from rav1d_b200 import lib
    lib.check(lib.init(0), "rb200_init")
    return lib
