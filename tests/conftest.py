import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _library_built():
    """The in-tree shared object must exist before any test imports it (nvcc cross-compiles without a GPU)."""
    from psydiff_b200 import build as b
    b.build_library()
