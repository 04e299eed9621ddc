import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference in oracle/_ref (skipped when absent)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    return oracle.lib()


@pytest.fixture(scope="session")
def gpu_ctx():
    import ogre_b200
    ctx = ogre_b200.Context(0)   # raises if libogrecu.so or the device is missing: no fallback
    yield ctx
    ctx.close()
