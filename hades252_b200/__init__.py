"""hades252_b200 -- B200-native batched Hades252 permutation engine (drop-in for the
`ScalarStrategy::perm` path of dusk-hades 0.24.1).  See DESIGN.md / INTEGRATION.md."""
from .strategy import (PARTIAL_ROUNDS, TOTAL_FULL_ROUNDS, WIDTH, CudaStrategy, HadesError,  # noqa: F401
                       Strategy)
from . import constants, scalar  # noqa: F401

__all__ = ["CudaStrategy", "Strategy", "HadesError", "WIDTH", "TOTAL_FULL_ROUNDS", "PARTIAL_ROUNDS", "constants", "scalar"]
