import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fe():
    """The ctypes binding of libplviwo_fe.so; builds the library if it is missing (nvcc cross-compiles on CPU)."""
    import plviwo_b200
    if not os.path.exists(plviwo_b200.LIB_PATH):
        plviwo_b200.build.build()
    return plviwo_b200


@pytest.fixture(scope="session")
def synth():
    import plviwo_b200
    return plviwo_b200.synth
