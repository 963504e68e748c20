import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference checkout at /root/reference (build container only)")


@pytest.fixture(scope="session")
def lib():
    from fempy_b200 import _build, _lib

    _build.build()
    return _lib.load()
