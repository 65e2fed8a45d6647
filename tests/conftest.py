import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and, if needed, the CUDA extension (nvcc cross-compiles without a GPU)."""
    from oracle import orc
    orc.build()
    from mc_water_ls_mw_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    yield
