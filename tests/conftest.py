import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import pyoracle
    return pyoracle.port()


@pytest.fixture(scope="session")
def oracle_best():
    """The real reference when its prebuilt library travelled with the repo, else the port."""
    from oracle import pyoracle
    return pyoracle.best()
