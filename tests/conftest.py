import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gpu():
    """One h264gpu context on cuda:0 through the C-ABI (fails loudly without it)."""
    import libh264_b200 as L
    g = L.Gpu(0)
    yield g
    g.close()
