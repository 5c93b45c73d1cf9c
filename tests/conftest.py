import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle (gcc) and the product library (nvcc, cross-compiles without a GPU) must
    exist before any test; both builds are no-ops when up to date."""
    from oracle import oracle
    oracle.build()
    from cupyimg_b200 import _build
    _build.build()
    yield
