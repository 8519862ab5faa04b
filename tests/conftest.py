import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")
for p in (ROOT, PKG_DIR, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kernels_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "kernels.npz"))


@pytest.fixture(scope="session")
def models_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "models.npz"))
