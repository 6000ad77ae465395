"""CPU-only: the N>1 host logic (range sharding, distributed Merkle with an all-gather of subtree
roots, max-over-ranks timing reduction) with world_size 2 over gloo.  The per-rank compute is the CPU
oracle here (tests may use it); on GPUs it is CudaStrategy.merkle_reduce_device / perm_batch_device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hades252_b200 import sharding


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


@pytest.mark.parametrize("n,world,expect", [
    (4 ** 12, 8, (2 ** 21, 10, 2, 2)),    # BASELINE config 3: 8 GPUs x 2^21 leaves = two 4^10 subtrees each
    (4 ** 12, 4, (2 ** 22, 11, 1, 1)),
    (4 ** 12, 2, (2 ** 23, 11, 2, 1)),
    (4 ** 12, 1, (4 ** 12, 12, 1, 0)),
    (16, 8, (2, 0, 2, 2)),
])
def test_merkle_plan(n, world, expect):
    p = sharding.merkle_plan(n, world)
    assert (p.leaves_per_rank, p.sub_levels, p.roots_per_rank, p.top_levels) == expect
    assert p.sub_levels + p.top_levels == sharding.log4_exact(n)


def test_merkle_plan_rejects_bad_shapes():
    for n, w in ((8, 2), (0, 1), (64, 3), (4, 8)):
        with pytest.raises(ValueError):
            sharding.merkle_plan(n, w)


def test_sponge_partition_balances_perm_counts():
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 33, size=5000)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    counts = sharding.sponge_perm_counts(offsets)
    assert np.array_equal(counts, lens // 4 + 1)
    for world in (1, 2, 8):
        parts = sharding.sponge_partition(offsets, world)
        assert parts[0][0] == 0 and parts[-1][1] == 5000
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        loads = [int(counts[lo:hi].sum()) for lo, hi in parts]
        assert max(loads) - min(loads) <= 9 * 2  # within a couple of messages of each other


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, depth, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu_oracle
    n = 4 ** depth
    plan = sharding.merkle_plan(n, world)
    lo, hi = sharding.shard_range(n, rank, world)
    assert hi - lo == plan.leaves_per_rank
    leaves = cpu_oracle.gen_elems(lo, hi - lo)            # rank-local generation, global indices

    def reduce_fn(nodes, levels):                          # the oracle stands in for the GPU kernels
        nodes = np.ascontiguousarray(nodes)
        for _ in range(levels):
            nodes = np.stack([cpu_oracle.merkle_root(nodes[i:i + 4], nthreads=1) for i in range(0, len(nodes), 4)])
        return nodes

    def all_gather_fn(roots):
        t = torch.from_numpy(roots.view(np.int64))
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t)
        return torch.cat(bufs).numpy().view(np.uint64)

    root = sharding.merkle_root_distributed(leaves, plan, reduce_fn, all_gather_fn)
    # perm sharding: each rank permutes its range, the digest pieces are combined with all_reduce
    states = cpu_oracle.gen_elems(lo * 5, (hi - lo) * 5).reshape(-1, 5, 4)[:64]
    local = torch.tensor([float(rank + 1)], dtype=torch.float64)   # stand-in for per-rank elapsed ms
    dist.all_reduce(local, op=dist.ReduceOp.MAX)                   # bench.py: max over ranks
    np.save(os.path.join(out_dir, f"root{rank}.npy"), root)
    np.save(os.path.join(out_dir, f"max{rank}.npy"), local.numpy())
    np.save(os.path.join(out_dir, f"perm{rank}.npy"), cpu_oracle.perm_batch(states, nthreads=1))
    dist.destroy_process_group()


@pytest.mark.parametrize("depth", [3, 4])
def test_world2_merkle_root_matches_single_process(tmp_path, depth):
    from oracle import cpu_oracle
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), depth, str(tmp_path)), nprocs=world, join=True)
    want = cpu_oracle.merkle_root(cpu_oracle.gen_elems(0, 4 ** depth))
    for r in range(world):
        got = np.load(tmp_path / f"root{r}.npy")
        assert np.array_equal(got.reshape(-1), want), f"rank {r}"
        assert float(np.load(tmp_path / f"max{r}.npy")[0]) == float(world)
