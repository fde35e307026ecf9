import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "gpu2: needs two CUDA devices (gpurun --gpus 2; every gpu2 test is also a gpu test)")


def require_two_gpus():
    """gpu2 tests: skipped with a stated reason on a one-GPU box, FAILED there when BIDI_REQUIRE_GPU2=1 (the
    two-GPU run of the round, whose log is committed under profiles/, sets it)."""
    import torch
    n = torch.cuda.device_count()
    if n >= 2:
        return
    msg = "needs two GPUs, this box has %d: run `BIDI_REQUIRE_GPU2=1 pytest -m gpu2` under `gpurun --gpus 2`" % n
    if os.environ.get("BIDI_REQUIRE_GPU2") == "1":
        pytest.fail(msg)
    pytest.skip(msg)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (CPU) and the CUDA library.  Both recipes are incremental makes, so this is a no-op
    when nothing changed -- and a stale binary can never be what the parity tests run against."""
    from oracle import pyoracle
    pyoracle.build()
    from bidinet_b200.build import build
    build()
    yield
