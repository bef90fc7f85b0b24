import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def rdx_paths():
    g = os.path.join(INPUTS, "init.rdx")
    return os.path.join(g, "input.xyz"), os.path.join(g, "ffield")


@pytest.fixture(scope="session")
def built():
    """Build the oracle and (cross-compile) the CUDA library once per session."""
    import __graft_entry__ as g
    g.build()
    return True
