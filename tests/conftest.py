import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (CPU) and make sure the CUDA library exists."""
    from oracle import pyoracle
    pyoracle.build()
    from bidinet_b200.build import LIB_PATH, build
    if not os.path.exists(LIB_PATH):
        build()
    yield
