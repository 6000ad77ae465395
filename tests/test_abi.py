"""CPU-only: the C-ABI library loads and exports every symbol include/hades_cuda.h declares; the
product fails loudly without a GPU; the product never touches oracle/."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hades_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hades_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from hades252_b200 import build
    build.build()
    from hades252_b200 import _native
    return _native.lib()


def test_header_symbols_all_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in hades_cuda.h but not exported"
    from hades252_b200 import _native
    assert sorted(_native.SIGNATURES) == declared


def test_library_is_sm100a_native():
    so = os.path.join(ROOT, "hades252_b200", "lib", "libhades_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_init_argument_errors(lib):
    from hades252_b200 import _native, constants
    ctx = _native.ctx_p()
    ark = np.ascontiguousarray(constants.round_constants())
    mds = np.ascontiguousarray(constants.mds_matrix(5))
    a, m = ark.ctypes.data_as(_native.u64p), mds.ctypes.data_as(_native.u64p)
    assert lib.hades_init(ctypes.byref(ctx), None, 1, 15, a, 960, m) == 1         # width out of range (2..14)
    assert b"width" in lib.hades_last_error(None)
    assert lib.hades_init(ctypes.byref(ctx), None, 1, 1, a, 960, m) == 1
    assert lib.hades_init(ctypes.byref(ctx), None, 1, 5, a, 100, m) == 6          # out of ARK constants
    assert b"out of ARK constants" in lib.hades_last_error(None)
    assert lib.hades_init(ctypes.byref(ctx), None, 0, 5, a, 960, m) == 1
    assert lib.hades_init(None, None, 1, 5, a, 960, m) == 1
    # null-context calls are rejected, not crashed
    assert lib.hades_perm_batch(None, None, 5) == 1
    assert lib.hades_width(None) == 0 and lib.hades_device_count(None) == 0
    lib.hades_destroy(None)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hades252_b200 import CudaStrategy, HadesError
    with pytest.raises(HadesError) as e:
        CudaStrategy([0])
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hades252_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|oracle/|libhades_oracle|hades_cpu\.c", text, flags=re.M):
                    if "tests/host_emul" in text and f.endswith(".cuh"):
                        continue
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_product_constants_match_reference_loader_semantics():
    """from_raw(bytes) == bytes*R mod p; first ROUND_CONSTANTS entry spot value."""
    from hades252_b200 import constants
    from oracle import hades_ref as H
    rc = constants.round_constants()
    assert rc.shape == (960, 4)
    blob = constants.ark_bin()
    for k in (0, 1, 334, 959):
        v = int.from_bytes(blob[32 * k:32 * k + 32], "little")
        assert [int(x) for x in rc[k]] == H.to_mont_limbs(v)
    for w in (3, 5, 9):
        m = constants.mds_matrix(w)
        assert m.shape == (w * w, 4)
        assert [int(x) for x in m[w + 1]] == H.to_mont_limbs(pow(1 + 1 + w, -1, H.P) * H.R % H.P)


def test_device_small_constants():
    """Montgomery forms of 1 and 15 hard-coded in kernels.cuh."""
    from oracle import hades_ref as H
    src = open(os.path.join(ROOT, "hades252_b200", "csrc", "width_impl.cuh")).read()
    for name, val in (("fr_set_one", 1), ("fr_set_fifteen", 15)):
        body = src[src.index(name):]
        words = re.findall(r"0x([0-9a-f]{8})u", body)[:8]
        got = sum(int(w, 16) << (32 * i) for i, w in enumerate(words))
        assert got == val * H.R % H.P, name


def test_device_merkle_bitmask_constants():
    """Montgomery forms of the bitmasks 1, 3, 7, 15 of partially filled Merkle nodes (kMaskMont)."""
    from oracle import hades_ref as H
    src = open(os.path.join(ROOT, "hades252_b200", "csrc", "width_impl.cuh")).read()
    body = src[src.index("kMaskMont[4][8]"):]
    words = re.findall(r"0x([0-9a-f]{8})u", body)[:32]
    for k in range(4):
        got = sum(int(w, 16) << (32 * i) for i, w in enumerate(words[8 * k:8 * k + 8]))
        assert got == ((1 << (k + 1)) - 1) * H.R % H.P


def test_merkle_tree_nodes_host_helper(lib):
    from oracle import cpu_oracle as C
    for n in (0, 1, 2, 4, 5, 16, 17, 1000, 4 ** 12, 4 ** 12 + 1):
        assert lib.hades_merkle_tree_nodes(n) == sum(C.merkle_level_sizes(n))
