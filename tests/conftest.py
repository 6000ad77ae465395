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


def _cuda_device_count() -> int:
    """CUDA devices visible to the C-ABI library's runtime; 0 when the driver / library is missing"""
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 0
    n = ctypes.c_int(0)
    return n.value if cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device: a plain `pytest tests` stays green
    there.  The driver's `-m gpu` run on the B200 box executes them all."""
    if not any("gpu" in item.keywords for item in items):
        return
    try:
        import torch
        have = torch.cuda.is_available() and torch.cuda.device_count() > 0
    except Exception:
        have = _cuda_device_count() > 0
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (the engine has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
