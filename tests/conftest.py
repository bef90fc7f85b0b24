import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "gpu2: needs two B200s on one node")


def _b200_count():
    """Number of sm_100 devices visible to this process (0 without a driver); asked of the CUDA runtime directly so that
    collecting the tests does not import torch."""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            rt = None
    if rt is None:
        return 0
    n = ctypes.c_int(0)
    if rt.cudaGetDeviceCount(ctypes.byref(n)) != 0:
        return 0
    good = 0
    for d in range(n.value):
        major = ctypes.c_int(0)
        if rt.cudaDeviceGetAttribute(ctypes.byref(major), 75, d) == 0 and major.value == 10:   # cudaDevAttrComputeCapabilityMajor
            good += 1
    return good


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a B200; gpu2-marked ones without a second B200."""
    ngpu = None
    for it in items:
        if "gpu" in it.keywords or "gpu2" in it.keywords:
            if ngpu is None:
                ngpu = _b200_count()
            if ngpu == 0:
                it.add_marker(pytest.mark.skip(reason="no sm_100 (B200) device: the CUDA hot path has no CPU fallback"))
            elif "gpu2" in it.keywords and ngpu < 2:
                it.add_marker(pytest.mark.skip(reason="needs two B200s"))


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
