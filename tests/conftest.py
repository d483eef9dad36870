import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle
    return pyoracle.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libspnngp_ref.so not built (needs /root/reference at build time)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def engine():
    from spnngp_b200 import Engine
    e = Engine()
    yield e
    e.close()
