import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Host library + CPU checkers are built once per session (seconds).  The CUDA library is built
    by __graft_entry__.build(); GPU tests fail loudly if it is missing."""
    from fastore_b200 import build
    build.build_host()
    build.build_oracle()
    yield
