import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context of the product library for the whole GPU session (fails loudly if the
    library or the device is missing -- there is no CPU fallback to fall through to)."""
    from typlonk_b200.ffi import Context
    c = Context(0)
    yield c
    c.close()
