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


@pytest.fixture(scope="session")
def seven_labels():
    return np.load(os.path.join(GOLDEN, "SevenLabels.npy"))


@pytest.fixture(scope="session")
def many_to_many():
    return np.load(os.path.join(GOLDEN, "ManyToMany.npy"))
