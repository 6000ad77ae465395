"""Multi-GPU host logic: one process per GPU, `torch.distributed` for the plumbing.

The permutation path shards trivially (independent states, no exchange): rank r owns the contiguous
range `shard_range(n, r, world)` and no collective touches the data.  The only exchange step in the
whole engine is the 4-ary Merkle tree: each rank reduces its leaf range to one or two subtree roots,
the roots are all-gathered (NCCL over NVLink on GPUs: 32-64 bytes per rank) and the few top levels
are finished redundantly on every rank.  

Nothing here computes a permutation: the reduction itself is injected (`reduce_fn`), which is
`CudaStrategy.merkle_reduce_device` on GPUs, and the CPU oracle in the world_size-2 gloo tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range of rank `rank`; same split as hades_perm_batch uses across devices."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n * rank // world, n * (rank + 1) // world


def log4_exact(n: int) -> int:
    """k such that n == 4^k, else -1."""
    if n < 1 or n & (n - 1):
        return -1
    lg = n.bit_length() - 1
    return lg // 2 if lg % 2 == 0 else -1


@dataclass(frozen=True)
class MerklePlan:
    n_leaves: int
    world: int
    leaves_per_rank: int
    sub_levels: int        # levels every rank reduces on its own range
    roots_per_rank: int    # 1 or 2
    top_levels: int        # levels left after the gather (on world * roots_per_rank nodes)

    @property
    def n_roots(self) -> int:
        return self.world * self.roots_per_rank


def merkle_plan(n_leaves: int, world: int) -> MerklePlan:
    """Split a 4^k-leaf tree over `world` ranks (world a power of two, each rank >= 1 subtree)."""
    depth = log4_exact(n_leaves)
    if depth < 0:
        raise ValueError("number of leaves must be a power of 4")
    if world < 1 or world & (world - 1):
        raise ValueError("world size must be a power of two")
    if n_leaves % world or n_leaves // world < 1:
        raise ValueError("more ranks than leaves")
    per = n_leaves // world                      # 4^a or 2 * 4^a
    sub_levels = 0
    while per % (4 ** (sub_levels + 1)) == 0:
        sub_levels += 1
    roots_per_rank = per // 4 ** sub_levels
    n_roots = roots_per_rank * world
    top = log4_exact(n_roots)
    if top < 0:
        # e.g. world = 2 with an odd split: fall back to one level less per rank
        while top < 0 and sub_levels > 0:
            sub_levels -= 1
            roots_per_rank = per // 4 ** sub_levels
            n_roots = roots_per_rank * world
            top = log4_exact(n_roots)
        if top < 0:
            raise ValueError("cannot split this tree over this many ranks")
    return MerklePlan(n_leaves, world, per, sub_levels, roots_per_rank, top)


def merkle_root_distributed(local_leaves, plan: MerklePlan, reduce_fn: Callable, all_gather_fn: Callable):
    """local_leaves: this rank's `plan.leaves_per_rank` leaves (any array type `reduce_fn` accepts).
    reduce_fn(nodes, levels) -> nodes / 4^levels ; all_gather_fn(roots) -> concatenation over ranks.
    Returns the root (one element) in the array type of reduce_fn."""
    roots = reduce_fn(local_leaves, plan.sub_levels)
    gathered = all_gather_fn(roots) if plan.world > 1 else roots
    return reduce_fn(gathered, plan.top_levels)


def sponge_perm_counts(offsets: np.ndarray) -> np.ndarray:
    """Permutations per message: floor(len / 4) + 1 (rate 4, padding always adds the single 1)."""
    lens = np.diff(np.asarray(offsets, dtype=np.uint64)).astype(np.int64)
    return lens // 4 + 1


def sponge_partition(offsets: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Message ranges per rank, contiguous and balanced by permutation count."""
    counts = sponge_perm_counts(offsets)
    n = counts.shape[0]
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, total * r / world, side="left")))
    bounds.append(n)
    bounds = [min(max(b, 0), n) for b in bounds]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world)]
