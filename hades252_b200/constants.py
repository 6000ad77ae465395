"""Host-side constant tables of the Hades252 instance, as the reference builds them at compile
time, regenerated from the published recipe instead of shipping binary blobs:

  * `round_constants()`  == `ROUND_CONSTANTS`  (reference: src/round_constants.rs:29-48, bytes of
    assets/ark.bin generated per assets/HOWTO.md:21-48)
  * `mds_matrix(width)`  == `MDS_MATRIX`       (reference: src/mds_matrix.rs:18-40, bytes of
    assets/mds.bin generated per assets/HOWTO.md:71-108; other widths per README.md:30-31)

Encoding quirk kept bit-for-bit: the asset files hold `internal_repr()` (Montgomery limbs) but the
loaders read them with `BlsScalar::from_raw`, which treats the bytes as a canonical integer and
converts to Montgomery form once more.  The tables returned here are the resulting IN-MEMORY limbs
(what `hades_init` uploads to `__constant__` memory).
"""
from __future__ import annotations

import hashlib
from functools import lru_cache

import numpy as np

MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001  # README.md:35
_R = (1 << 256) % MODULUS
N_ROUND_CONSTANTS = 960  # round_constants.rs:16

ARK_BIN_SHA256 = "78c427449282315729eaa2e39e1937e0aa0b010c4c38bcbb1d57016011880485"
MDS_BIN_SHA256 = {5: "131915cbeae1bde75422cce7fcf7feb9223a4dec370a937a2133c1f998ded0e7"}


def ark_bin() -> bytes:
    """Bytes of assets/ark.bin (HOWTO.md:21-48)."""
    h, prev, out = b"poseidon-for-plonk", 1, bytearray()
    for _ in range(N_ROUND_CONSTANTS):
        h = hashlib.sha512(h).digest()
        prev = (int.from_bytes(h, "little") + prev) % MODULUS  # from_bytes_wide(h) + p
        out += (prev * _R % MODULUS).to_bytes(32, "little")    # internal_repr()
    return bytes(out)


def mds_bin(width: int) -> bytes:
    """Bytes of assets/mds.bin for `width` (HOWTO.md:71-108): Cauchy 1/(i + j + width)."""
    out = bytearray()
    for i in range(width):
        for j in range(width):
            out += (pow(i + j + width, -1, MODULUS) * _R % MODULUS).to_bytes(32, "little")
    return bytes(out)


def from_raw_table(blob: bytes) -> np.ndarray:
    """The loaders' `BlsScalar::from_raw([a,b,c,d])` over a blob of 32-byte LE entries
    (round_constants.rs:36-41, lib.rs:33-44): returns uint64 [n,4] Montgomery limbs."""
    n = len(blob) // 32
    out = np.empty((n, 4), dtype=np.uint64)
    for k in range(n):
        v = int.from_bytes(blob[32 * k:32 * k + 32], "little")
        if v >= MODULUS:
            raise ValueError("asset entry is not a canonical field element")
        m = v * _R % MODULUS
        for i in range(4):
            out[k, i] = (m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
    return out


@lru_cache(maxsize=None)
def _round_constants() -> np.ndarray:
    blob = ark_bin()
    assert hashlib.sha256(blob).hexdigest() == ARK_BIN_SHA256, "regenerated ark.bin differs from the reference asset"
    t = from_raw_table(blob)
    t.setflags(write=False)
    return t


@lru_cache(maxsize=None)
def _mds_matrix(width: int) -> np.ndarray:
    blob = mds_bin(width)
    if width in MDS_BIN_SHA256:
        assert hashlib.sha256(blob).hexdigest() == MDS_BIN_SHA256[width], "regenerated mds.bin differs from the reference asset"
    t = from_raw_table(blob)
    t.setflags(write=False)
    return t


def round_constants() -> np.ndarray:
    """uint64 [960,4]: in-memory limbs of ROUND_CONSTANTS."""
    return _round_constants()


def mds_matrix(width: int = 5) -> np.ndarray:
    """uint64 [width*width,4]: in-memory limbs of MDS_MATRIX, row-major."""
    return _mds_matrix(int(width))
