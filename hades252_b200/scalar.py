"""Representation helpers for `BlsScalar` values: integer <-> the 4 little-endian u64 Montgomery limbs the
engine (and the reference's `BlsScalar`) holds in memory.  Counterparts of `BlsScalar::from(u64)` /
`from_raw` / canonical reduction used by the reference's tests (src/strategies/scalar.rs:64-66).
Pure representation changes with Python integers: nothing here evaluates the permutation."""
from __future__ import annotations

from typing import Iterable, List

import numpy as np

from .constants import MODULUS

_R = (1 << 256) % MODULUS
_R_INV = pow(_R, -1, MODULUS)
_MASK = (1 << 64) - 1


def from_int(x: int) -> np.ndarray:
    """canonical integer (any size, reduced mod p) -> uint64[4] Montgomery limbs."""
    m = (x % MODULUS) * _R % MODULUS
    return np.array([(m >> (64 * i)) & _MASK for i in range(4)], dtype=np.uint64)


def to_int(limbs) -> int:
    """uint64[4] Montgomery limbs -> canonical integer in [0, p)."""
    m = sum(int(l) << (64 * i) for i, l in enumerate(np.asarray(limbs, dtype=np.uint64).reshape(4)))
    if m >= MODULUS:
        raise ValueError("limbs are not a fully reduced BlsScalar")
    return m * _R_INV % MODULUS


def state_from_ints(values: Iterable[int]) -> np.ndarray:
    """[WIDTH] integers -> uint64[WIDTH, 4], ready for `CudaStrategy.perm`."""
    return np.stack([from_int(v) for v in values])


def state_to_ints(state) -> List[int]:
    return [to_int(w) for w in np.asarray(state, dtype=np.uint64).reshape(-1, 4)]
