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


@pytest.fixture
def tmp_path_nodash():
    """A scratch directory WITHOUT '-' in its path: like the reference (src/main.c:128,136), the CLI
    takes any -i/-o argument that contains a '-' to mean stdin/stdout (pytest's own tmp_path has one)."""
    import shutil
    import tempfile
    from pathlib import Path
    d = tempfile.mkdtemp(prefix="fmd_", dir="/tmp")
    assert "-" not in d
    yield Path(d)
    shutil.rmtree(d, ignore_errors=True)
