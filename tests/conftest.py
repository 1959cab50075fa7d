import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libavcer_b200.so")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return {name: np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
            for name in ("fusion", "preprocess", "video", "audio", "weight_search", "resample", "pred_av", "face")}


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library on a real device; GPU tests must fail loudly if it is missing."""
    from avcer_b200 import _lib

    _lib.load()
    _lib.require_device()
    return _lib
