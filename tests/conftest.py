import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "hades252_kat.json")) as f:
        return json.load(f)


def limbs_to_array(limbs_hex):
    """[[hex x4] x W] -> uint64 [W,4]"""
    return np.array([[int(x, 16) for x in word] for word in limbs_hex], dtype=np.uint64)


@pytest.fixture(scope="session")
def cuda_strategy():
    from hades252_b200 import CudaStrategy
    s = CudaStrategy([0])
    yield s
    s.close()
