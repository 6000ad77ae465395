"""CPU-only: the Rust side staged for the day a toolchain exists (bindings/rust) stays consistent with the committed
golden vectors and with the reference crate's files.  Nothing here compiles Rust (none in this image)."""
import json
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "bindings", "rust")


def test_kat_rs_expects_the_committed_golden_vectors(golden):
    src = open(os.path.join(RUST, "tests", "kat.rs")).read()
    by = {c["name"]: c for c in golden["perm"]}
    kats = re.findall(r"Kat \{ input: \[([^\]]+)\], out0: \[([^\]]+)\], out4: \[([^\]]+)\] \}", src)
    assert len(kats) == 4
    names = {"1, 1, 1, 1, 1": "readme_example_ones", "17, 17, 17, 17, 17": "hades_det_17",
             "5000, 5000, 5000, 5000, 5000": "preimage_constant_5000", "0, 1, 2, 3, 4": "iota"}
    for inp, out0, out4 in kats:
        c = by[names[inp]]
        assert [int(x, 16) for x in out0.split(",")] == [int(x, 16) for x in c["output_mont_limbs"][0]]
        assert [int(x, 16) for x in out4.split(",")] == [int(x, 16) for x in c["output_mont_limbs"][4]]
    a = re.search(r'READING_A_ONES_WORD0: &str = "([0-9a-f]+)"', src).group(1)
    assert int(a, 16) == int(by["readme_example_ones"]["output"][0], 16)
    b = re.search(r'READING_B_ONES_WORD0: &str = "([0-9a-f]+)"', src).group(1)
    assert a != b


def test_ffi_rs_binds_only_declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "hades_cuda.h")).read()
    ffi = open(os.path.join(RUST, "src", "strategies", "ffi.rs")).read()
    bound = re.findall(r"pub fn (hades_[a-z0-9_]+)\(", ffi)
    assert len(bound) >= 12
    for name in bound:
        assert re.search(r"\b%s\s*\(" % name, hdr), f"{name} bound in ffi.rs but not declared in hades_cuda.h"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present on this box")
def test_overlay_patch_applies_to_the_reference(tmp_path):
    dst = tmp_path / "crate"
    shutil.copytree("/root/reference", dst, ignore=shutil.ignore_patterns(".git"))
    patch = os.path.join(RUST, "patches", "cuda_feature.patch")
    res = subprocess.run(["patch", "-p1", "-i", patch], cwd=dst, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert 'cuda = []' in open(dst / "Cargo.toml").read()
    s = open(dst / "src" / "strategies.rs").read()
    assert "mod cuda;" in s and "pub use cuda::{CudaError, CudaStrategy};" in s
    assert "pub use strategies::{CudaError, CudaStrategy};" in open(dst / "src" / "lib.rs").read()
