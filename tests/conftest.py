import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_ctx():
    """A lash_ctx on cuda:0.  Fails loudly (no skip) when the CUDA library or the device is missing:
    a silent skip would hide a CPU fallback."""
    from lash_b200.ops import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
