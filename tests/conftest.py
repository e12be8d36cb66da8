import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The built shared library; building is cheap when up to date (make)."""
    from vampomi_b200 import build, capi
    if not os.path.isfile(build.LIB_PATH):
        build.build()
    return capi.load_library()
