import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Build (or reuse) the in-tree libraries: the CUDA library cross-compiles without a GPU."""
    from sybil_b200 import _build
    _build.build_all()


@pytest.fixture(autouse=True)
def reset_flags():
    from sybil_b200.engine import FLAGS
    FLAGS.reset()
    yield
    FLAGS.reset()
