import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def _cuda_device_usable():
    """True when pq_create succeeds on device 0.  A missing library is NOT a reason to skip:
    the GPU tests must then fail loudly (no CPU fallback), so only 'library loads but there is
    no usable device' returns False."""
    import ctypes
    lib_path = os.path.join(ROOT, "picoquant.jl_b200", "csrc", "libpq_b200.so")
    if not os.path.isfile(lib_path):
        return True
    try:
        lib = ctypes.CDLL(lib_path)
        lib.pq_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.pq_destroy.argtypes = [ctypes.c_void_p]
        h = ctypes.c_void_p()
        rc = lib.pq_create(0, 1, ctypes.byref(h))
        if rc == 0:
            lib.pq_destroy(h)
        return rc == 0
    except OSError:
        return True


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _cuda_device_usable():
        skip = pytest.mark.skip(reason="no usable CUDA device (pq_create failed)")
        for it in gpu_items:
            it.add_marker(skip)
