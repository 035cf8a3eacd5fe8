import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One tmc2gpu context per test session (GPU tests only)."""
    import tmc2rs_b200  # noqa: F401
    from tmc2rs_b200 import codec
    ctx = codec.Context()
    yield ctx
    ctx.close()
