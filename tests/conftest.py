import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    """one probe per session, through the product's own entry point (mci_create fails without a usable device)"""
    try:
        import ctypes
        from itkcontourinterpolation_b200 import _lib as L
        lib = L.lib()
        ctx = ctypes.c_void_p()
        if lib.mci_create(ctypes.byref(ctx), 0) != 0:
            return False
        lib.mci_destroy(ctx)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (the product has no CPU fallback; run with -m gpu on a B200)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def seven_labels():
    return np.load(os.path.join(GOLDEN, "SevenLabels.npy"))


@pytest.fixture(scope="session")
def many_to_many():
    return np.load(os.path.join(GOLDEN, "ManyToMany.npy"))
