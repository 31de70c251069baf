import os
import sys

import pytest

# Several in-process ranks of a data-parallel group on ONE device (tests/test_gpu_dp.py) wait
# for each other on the device: their streams must not share a hardware queue.  Must be set
# before the first CUDA call of the process.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import ctypes
        from plasticity_b200 import _lib
        n = ctypes.c_int(0)
        return _lib.load().pb_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (oracle) once; the product library must already be
    built by __graft_entry__.build() — tests never build or patch the product."""
    from oracle import pyoracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        pyoracle.build()
    yield
