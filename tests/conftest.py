import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The shared libraries are built in-tree by __graft_entry__.build(); build them if missing."""
    import __graft_entry__ as ge
    ge.build(quiet=True)
