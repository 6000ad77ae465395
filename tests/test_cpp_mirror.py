"""The C++ host mirror (include/hades_strategy.hpp) compiled with g++ against the C-ABI library.
CPU box: it must build, link, and REFUSE to run without a GPU (exit 77, no CPU fallback).
GPU box (-m gpu): hades_det + known answer + perm_batch through the C++ CudaStrategy."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "hades252_b200", "lib")


def _build(tmp_path):
    from hades252_b200 import build
    build.build()
    exe = str(tmp_path / "strategy_main")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "strategy_main.cpp"),
                           "-L" + LIBDIR, "-lhades_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def _args(golden):
    from oracle import hades_ref as H
    case = next(c for c in golden["perm"] if c["name"] == "hades_det_17")
    limbs = H.to_mont_limbs(17) + H.to_mont_limbs(19) + [int(x, 16) for x in case["output_mont_limbs"][0]]
    return ["%x" % l for l in limbs]


def test_generated_constants_header_matches_python_tables():
    import re
    from hades252_b200 import constants
    text = open(os.path.join(ROOT, "include", "hades_constants.h")).read()
    vals = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", text)]
    want = np.concatenate([constants.round_constants().reshape(-1)] +
                          [constants.mds_matrix(w).reshape(-1) for w in (3, 5, 9)])
    assert vals == [int(x) for x in want]


def test_cpp_mirror_builds_and_refuses_without_gpu(tmp_path, golden):
    import torch
    exe = _build(tmp_path)
    res = subprocess.run([exe, *_args(golden)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert res.returncode == 0, res.stdout + res.stderr
    else:
        assert res.returncode == 77 and "no CPU fallback" in res.stdout, res.stdout + res.stderr


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path, golden):
    exe = _build(tmp_path)
    res = subprocess.run([exe, *_args(golden)], capture_output=True, text=True)
    assert res.returncode == 0 and "cpp strategy OK" in res.stdout, res.stdout + res.stderr


def test_c_example_builds(tmp_path):
    """examples/perm_batch.c compiles and links against the C ABI with a plain C compiler."""
    from hades252_b200 import build
    build.build()
    exe = str(tmp_path / "perm_batch")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-o", exe, os.path.join(ROOT, "examples", "perm_batch.c"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + LIBDIR, "-lhades_b200", "-Wl,-rpath," + LIBDIR])
    import torch
    res = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        # known answer: SURVEY.md 8(c), Montgomery limbs of perm([1;5])[0]
        assert res.returncode == 0 and "935feb66a5e6cf3c 2409c7dd1a61ab1c 832c33cbf2dd481f 23338e018f505a2a" in res.stdout, res.stdout
    else:
        assert res.returncode == 4 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_c_example_on_gpu(tmp_path):
    test_c_example_builds(tmp_path)


def test_scalar_helpers_round_trip():
    from hades252_b200 import scalar
    from oracle import hades_ref as H
    for v in (0, 1, 17, 5000, H.P - 1, H.P + 5, 1 << 255):
        limbs = scalar.from_int(v)
        assert [int(x) for x in limbs] == H.to_mont_limbs(v % H.P)
        assert scalar.to_int(limbs) == v % H.P
    st = scalar.state_from_ints([1, 2, 3, 4, 5])
    assert st.shape == (5, 4) and scalar.state_to_ints(st) == [1, 2, 3, 4, 5]
    with pytest.raises(ValueError):
        scalar.to_int(np.array([2**64 - 1] * 4, dtype=np.uint64))
