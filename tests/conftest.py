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


@pytest.fixture(scope="session")
def config_db():
    """BASELINE.json's seeded databases (tests/golden/config_digests.json: generator arguments), generated once per session."""
    import json
    from yacht_b200 import synth
    with open(os.path.join(ROOT, "tests", "golden", "config_digests.json")) as f:
        digests = json.load(f)
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = synth.make_reference_db(**digests[name]["generator"])
        return cache[name]
    return get
