import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the in-tree libraries once if they are missing (CPU-only build works: nvcc
    cross-compiles sm_100a without a GPU)."""
    import __graft_entry__ as g
    g.build(only_missing=True)
