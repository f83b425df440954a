import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One context on cuda:0 through the C ABI.  Fails loudly when the extension or GPU is missing."""
    from yacht_b200 import _lib
    ctx = _lib.GpuContext(0)
    yield ctx
    ctx.close()
