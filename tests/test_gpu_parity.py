"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on identical inputs.  Bit-exact everywhere: integer work."""
import numpy as np
import pytest

from conftest import limbs_to_array

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    from oracle import cpu_oracle
    return cpu_oracle


@pytest.fixture(scope="module")
def H():
    from oracle import hades_ref
    return hades_ref


def test_extension_loaded_and_kernels_spill_free(cuda_strategy):
    info = cuda_strategy.kernel_info("perm")
    assert info["regs_per_thread"] > 0
    print("perm5", info, "merkle", cuda_strategy.kernel_info("merkle"), "sponge", cuda_strategy.kernel_info("sponge"))


@pytest.mark.parametrize("algo,regs", [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (2, 0), (2, 3), (2, 6), (2, 9)])
def test_all_kernel_variants_bit_identical(oracle, algo, regs):
    """dense schedule (reference round structure) vs sparse-partial-round schedule, all register
    budgets: same bits as the oracle, for perm, merkle and sponge."""
    from hades252_b200 import CudaStrategy
    n = 3000
    s = oracle.gen_elems(321, 5 * n).reshape(n, 5, 4)
    want = oracle.perm_batch(s)
    with CudaStrategy([0]) as strat:
        strat.set_variant(algo, regs)
        got = s.copy()
        strat.perm_batch(got)
        assert np.array_equal(got, want)
        leaves = oracle.gen_elems(77, 4 ** 5)
        assert np.array_equal(strat.merkle_root(leaves), oracle.merkle_root(leaves))
        lens = np.arange(200) % 13
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        elems = oracle.gen_elems(5, int(offsets[-1]))
        assert np.array_equal(strat.sponge_batch(elems, offsets), oracle.sponge_batch(elems, offsets))
        print(algo, regs, strat.kernel_info("perm"))


@pytest.mark.parametrize("w", [3, 9])
@pytest.mark.parametrize("algo", [0, 1, 2])
def test_other_widths_both_schedules(oracle, w, algo):
    from hades252_b200 import CudaStrategy
    n = 2000
    s = oracle.gen_elems(55, w * n).reshape(n, w, 4)
    with CudaStrategy([0], width=w) as strat:
        strat.set_variant(algo, 2 if w == 9 else 0)
        got = s.copy()
        strat.perm_batch(got)
    assert np.array_equal(got, oracle.perm_batch(s, w))


def test_golden_vectors(cuda_strategy, golden):
    from hades252_b200 import CudaStrategy
    strategies = {5: cuda_strategy}
    try:
        for c in golden["perm"]:
            w = c["width"]
            if w not in strategies:
                strategies[w] = CudaStrategy([0], width=w)
            state = limbs_to_array(c["input_mont_limbs"]).copy()
            strategies[w].perm(state)
            assert np.array_equal(state, limbs_to_array(c["output_mont_limbs"])), c["name"]
    finally:
        for w, s in strategies.items():
            if w != 5:
                s.close()


def test_reference_self_consistency_tests(cuda_strategy, H):
    """scalar.rs:62-74 hades_det and README.md:50-65, through the device strategy."""
    def st(v):
        return np.array([H.to_mont_limbs(v)] * 5, dtype=np.uint64)
    x, y, z = st(17), st(17), st(19)
    for a in (x, y, z):
        cuda_strategy.perm(a)
    assert np.array_equal(x, y) and not np.array_equal(x, z)
    one = st(1)
    out = one.copy()
    cuda_strategy.perm(out)
    assert not np.array_equal(out, one) and out.shape == one.shape


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 128, 129, 1000, 4097])
def test_ragged_batch_sizes(cuda_strategy, oracle, n):
    s = oracle.gen_elems(n * 17, 5 * n).reshape(n, 5, 4)
    got = s.copy()
    cuda_strategy.perm_batch(got)
    assert np.array_equal(got, oracle.perm_batch(s))


def test_empty_batch(cuda_strategy):
    s = np.empty((0, 5, 4), dtype=np.uint64)
    cuda_strategy.perm_batch(s)


def test_wrong_shape_rejected(cuda_strategy):
    with pytest.raises(ValueError):
        cuda_strategy.perm(np.zeros((4, 4), dtype=np.uint64))
    with pytest.raises(ValueError):
        cuda_strategy.perm_batch(np.zeros((3, 4, 4), dtype=np.uint64))
    with pytest.raises(TypeError):
        cuda_strategy.perm_batch(np.zeros((3, 5, 4), dtype=np.int64))


def test_edge_values(cuda_strategy, oracle, H):
    P = H.P
    vals = [0, 1, 2, P - 1, P - 2, H.R, H.R2, (1 << 255) % P, (1 << 254), 0xFFFFFFFF, (1 << 64) - 1, (1 << 128) - 1,
            P >> 1, (P >> 1) + 1]
    rng = np.random.default_rng(5)
    states = []
    for _ in range(256):
        pick = rng.integers(0, len(vals), size=5)
        states.append([H.to_mont_limbs(vals[k]) for k in pick])
    # raw limb patterns that force carries / final subtractions: p-1 as LIMBS, ff..ff low limbs
    pm1 = [(P - 1 >> (64 * i)) & (2**64 - 1) for i in range(4)]
    ffs = [2**64 - 1, 2**64 - 1, 2**64 - 1, 0x73eda753299d7d47]
    states.append([pm1] * 5)
    states.append([ffs] * 5)
    states.append([pm1, ffs, [0, 0, 0, 0], [1, 0, 0, 0], ffs])
    s = np.array(states, dtype=np.uint64)
    got = s.copy()
    cuda_strategy.perm_batch(got)
    assert np.array_equal(got, oracle.perm_batch(s))


def test_config1_2pow20_states_full_compare(cuda_strategy, oracle):
    """BASELINE.json configs[0]: 2^20 random width-5 states, every output limb compared."""
    n = 1 << 20
    s = oracle.gen_elems(0, 5 * n).reshape(n, 5, 4)
    got = s.copy()
    cuda_strategy.perm_batch(got)
    want = oracle.perm_batch(s)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("w", [3, 9])
def test_other_widths(oracle, w):
    from hades252_b200 import CudaStrategy
    n = 5000
    s = oracle.gen_elems(99, w * n).reshape(n, w, 4)
    with CudaStrategy([0], width=w) as strat:
        got = s.copy()
        strat.perm_batch(got)
    assert np.array_equal(got, oracle.perm_batch(s, w))


def test_unsupported_width_fails_loudly():
    from hades252_b200 import CudaStrategy, HadesError
    with pytest.raises(HadesError):
        CudaStrategy([0], width=15)   # 67 * 15 > 960 round constants
    with pytest.raises(HadesError):
        CudaStrategy([0], width=1)
    with pytest.raises(HadesError):
        CudaStrategy([99])


def test_device_pointer_api_and_idempotent_layout(cuda_strategy, oracle):
    import torch
    n = 10000
    s = oracle.gen_elems(4242, 5 * n).reshape(n, 5, 4)
    d = torch.from_numpy(s.view(np.int64)).cuda()
    stream = torch.cuda.current_stream()
    cuda_strategy.perm_batch_device(d.data_ptr(), n, stream.cuda_stream)
    stream.synchronize()
    got = d.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, oracle.perm_batch(s))


def test_device_generator_and_digest_match_oracle(cuda_strategy, oracle):
    import torch
    n_elems = 5 * 3000
    d = torch.empty(n_elems * 4, dtype=torch.int64, device="cuda")
    cuda_strategy.gen_elems_device(d.data_ptr(), 777, n_elems, 0x4861646573323532)
    dig = torch.zeros(4, dtype=torch.int64, device="cuda")
    cuda_strategy.digest_device(d.data_ptr(), 0, n_elems * 4, dig.data_ptr())
    torch.cuda.synchronize()
    host = d.cpu().numpy().view(np.uint64).reshape(n_elems, 4)
    assert np.array_equal(host, oracle.gen_elems(777, n_elems))
    assert np.array_equal(dig.cpu().numpy().view(np.uint64), oracle.digest(host))


@pytest.mark.parametrize("depth", [0, 1, 2, 3, 5, 8])
def test_merkle_root(cuda_strategy, oracle, depth):
    n = 4 ** depth
    leaves = oracle.gen_elems(31337, n)
    assert np.array_equal(cuda_strategy.merkle_root(leaves), oracle.merkle_root(leaves))


def test_merkle_golden(cuda_strategy, golden, H):
    for m in golden["merkle"]:
        leaves = np.array([H.to_mont_limbs(i) for i in range(m["leaves"])], dtype=np.uint64)
        assert [int(x) for x in cuda_strategy.merkle_root(leaves)] == [int(l, 16) for l in m["root_mont_limbs"]]


def test_merkle_rejects_non_power_of_4(cuda_strategy):
    from hades252_b200 import HadesError
    for n in (0, 2, 8, 12, 32):
        with pytest.raises(HadesError) as e:
            cuda_strategy.merkle_root(np.zeros((n, 4), dtype=np.uint64))
        assert e.value.status in (1, 2)


def test_merkle_sharded_equals_single(cuda_strategy, oracle):
    """The multi-GPU decomposition on one GPU: reduce 8 'virtual shards' to subtree roots with the
    device API, then reduce the gathered roots -- must equal the one-shot root."""
    import torch
    depth, shards = 7, 8
    n = 4 ** depth
    leaves = oracle.gen_elems(5, n)
    want = oracle.merkle_root(leaves)
    per = n // shards                      # 2 * 4^5
    sub_levels = 5
    roots = []
    for g in range(shards):
        d = torch.from_numpy(leaves[g * per:(g + 1) * per].view(np.int64).copy()).cuda()
        scratch = torch.empty((per // 4 + per // 16 + 4) * 4, dtype=torch.int64, device="cuda")
        out = torch.empty((per >> (2 * sub_levels)) * 4, dtype=torch.int64, device="cuda")
        cuda_strategy.merkle_reduce_device(d.data_ptr(), per, sub_levels, scratch.data_ptr(), out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
        roots.append(out)
    allr = torch.cat(roots)                # 16 roots
    scratch = torch.empty(64 * 4, dtype=torch.int64, device="cuda")
    out = torch.empty(4, dtype=torch.int64, device="cuda")
    cuda_strategy.merkle_reduce_device(allr.data_ptr(), 16, 2, scratch.data_ptr(), out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)


def test_sponge_golden(cuda_strategy, golden, H):
    msgs = [[int(x, 16) for x in s["message"]] for s in golden["sponge"]]
    elems = np.array([H.to_mont_limbs(x) for m in msgs for x in m], dtype=np.uint64).reshape(-1, 4)
    offsets = np.cumsum([0] + [len(m) for m in msgs]).astype(np.uint64)
    dig = cuda_strategy.sponge_batch(elems, offsets)
    for k, s in enumerate(golden["sponge"]):
        assert [int(x) for x in dig[k]] == [int(l, 16) for l in s["digest_mont_limbs"]], s["message"]


def test_sponge_variable_lengths(cuda_strategy, oracle):
    """config 4 shape at reduced size: lengths 1 + (splitmix64 mod 32), plus empty messages."""
    rng = np.random.default_rng(11)
    n = 20000
    lens = rng.integers(0, 33, size=n)
    lens[:7] = [0, 1, 3, 4, 5, 8, 32]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    elems = oracle.gen_elems(9, int(offsets[-1]))
    got = cuda_strategy.sponge_batch(elems, offsets)
    assert np.array_equal(got, oracle.sponge_batch(elems, offsets))


def test_sponge_all_empty_and_zero_messages(cuda_strategy, oracle):
    offsets = np.zeros(6, dtype=np.uint64)
    got = cuda_strategy.sponge_batch(np.empty((0, 4), dtype=np.uint64), offsets)
    assert np.array_equal(got, oracle.sponge_batch(np.empty((0, 4), dtype=np.uint64), offsets))
    assert cuda_strategy.sponge_batch(np.empty((0, 4), dtype=np.uint64), np.zeros(1, dtype=np.uint64)).shape == (0, 4)


def test_multi_context_and_virtual_shards(oracle):
    """Sharding logic with the same device listed twice (two 'virtual' GPUs)."""
    from hades252_b200 import CudaStrategy
    n = 3001
    s = oracle.gen_elems(1, 5 * n).reshape(n, 5, 4)
    with CudaStrategy([0, 0]) as two:
        got = s.copy()
        two.perm_batch(got)
        assert np.array_equal(got, oracle.perm_batch(s))
        leaves = oracle.gen_elems(2, 4 ** 6)
        assert np.array_equal(two.merkle_root(leaves), oracle.merkle_root(leaves))


def test_large_batch_digest_property(cuda_strategy, oracle):
    """2^22 states generated and permuted on device; a strided 2^12 sample is checked against the
    oracle and the digest of the outputs must be reproducible (determinism at full size)."""
    import torch
    n = 1 << 22
    d = torch.empty(n * 20, dtype=torch.int64, device="cuda")
    cuda_strategy.gen_elems_device(d.data_ptr(), 0, n * 5, 0x4861646573323532)
    idx = torch.arange(0, n, n >> 12, device="cuda")
    before = d.view(n, 20)[idx].cpu().numpy().view(np.uint64).reshape(-1, 5, 4)
    cuda_strategy.perm_batch_device(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    after = d.view(n, 20)[idx].cpu().numpy().view(np.uint64).reshape(-1, 5, 4)
    assert np.array_equal(after, oracle.perm_batch(before))
    digs = []
    for _ in range(2):
        cuda_strategy.gen_elems_device(d.data_ptr(), 0, n * 5, 0x4861646573323532)
        cuda_strategy.perm_batch_device(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        dig = torch.zeros(4, dtype=torch.int64, device="cuda")
        cuda_strategy.digest_device(d.data_ptr(), 0, n * 20, dig.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        digs.append(dig.cpu().numpy().copy())
    assert np.array_equal(digs[0], digs[1])


def test_cuda_graph_capture_of_device_calls(cuda_strategy, oracle):
    """The device-resident entry points only enqueue work on the caller's stream, so a whole Merkle
    reduction (6 level launches) and a perm_batch can be captured in a CUDA graph and replayed."""
    import torch
    n = 4 ** 6
    leaves = oracle.gen_elems(99, n)
    d = torch.from_numpy(leaves.view(np.int64).copy()).cuda()
    scratch = torch.empty((n // 4 + n // 16 + 4) * 4, dtype=torch.int64, device="cuda")
    out = torch.zeros(4, dtype=torch.int64, device="cuda")
    states = torch.from_numpy(oracle.gen_elems(5, 5 * 512).view(np.int64).copy()).cuda()
    s0 = states.clone()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        cuda_strategy.merkle_reduce_device(d.data_ptr(), n, 6, scratch.data_ptr(), out.data_ptr(), side.cuda_stream)  # warm-up
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            cuda_strategy.merkle_reduce_device(d.data_ptr(), n, 6, scratch.data_ptr(), out.data_ptr(), side.cuda_stream)
            cuda_strategy.perm_batch_device(states.data_ptr(), 512, side.cuda_stream)
    out.zero_()
    states.copy_(s0)
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), oracle.merkle_root(leaves))
    want = oracle.perm_batch(s0.cpu().numpy().view(np.uint64).reshape(512, 5, 4))
    assert np.array_equal(states.cpu().numpy().view(np.uint64).reshape(512, 5, 4), want)
    g.replay()  # second replay permutes the permuted states again
    torch.cuda.synchronize()
    assert np.array_equal(states.cpu().numpy().view(np.uint64).reshape(512, 5, 4), oracle.perm_batch(want))


def test_host_register_path(cuda_strategy, oracle):
    n = 50000
    s = oracle.gen_elems(8, 5 * n).reshape(n, 5, 4)
    got = s.copy()
    cuda_strategy.host_register(got.ctypes.data, got.nbytes)
    try:
        cuda_strategy.perm_batch(got)
    finally:
        cuda_strategy.host_unregister(got.ctypes.data)
    assert np.array_equal(got, oracle.perm_batch(s))


def test_concurrent_contexts_from_threads(oracle):
    """One context per caller thread (the `&mut self` contract); contexts may run concurrently."""
    import threading
    from hades252_b200 import CudaStrategy
    n = 20000
    inputs = [oracle.gen_elems(1000 * t, 5 * n).reshape(n, 5, 4) for t in range(4)]
    outs = [None] * 4

    def work(t):
        with CudaStrategy([0]) as strat:
            o = inputs[t].copy()
            for _ in range(2):
                strat.perm_batch(o)
            outs[t] = o

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for t in range(4):
        assert np.array_equal(outs[t], oracle.perm_batch(oracle.perm_batch(inputs[t])))


def test_permutation_is_a_bijection_on_a_sample(cuda_strategy, oracle):
    """Size-independent property: distinct inputs give distinct outputs (2^16 states)."""
    n = 1 << 16
    s = oracle.gen_elems(31, 5 * n).reshape(n, 5, 4)
    o = s.copy()
    cuda_strategy.perm_batch(o)
    assert len({bytes(x) for x in o.reshape(n, -1)}) == n


def test_single_process_multi_device_context(oracle):
    """`CudaStrategy::new(&[0, 1, ...])`: one context over every GPU of the box; host batches, Merkle
    leaves and sponge messages are sharded inside the C library (skipped on a 1-GPU box)."""
    import torch
    from hades252_b200 import CudaStrategy
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 300001
    s = oracle.gen_elems(11, 5 * n).reshape(n, 5, 4)
    with CudaStrategy(list(range(g))) as strat:
        got = s.copy()
        strat.perm_batch(got)
        assert np.array_equal(got, oracle.perm_batch(s))
        leaves = oracle.gen_elems(12, 4 ** 8)
        assert np.array_equal(strat.merkle_root(leaves), oracle.merkle_root(leaves))
        lens = np.random.default_rng(2).integers(0, 33, size=60000)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        elems = oracle.gen_elems(13, int(offsets[-1]))
        assert np.array_equal(strat.sponge_batch(elems, offsets), oracle.sponge_batch(elems, offsets))


@pytest.mark.parametrize("w", [2, 4, 6, 7, 8, 10, 14])
def test_generic_width_kernel(oracle, w):
    """Widths without a tuned kernel (the reference allows any WIDTH with 67*WIDTH <= 960, README.md:30-31)
    run the generic kernel; assets regenerated per assets/HOWTO.md."""
    from hades252_b200 import CudaStrategy, HadesError
    n = 700
    s = oracle.gen_elems(w * 100, w * n).reshape(n, w, 4)
    with CudaStrategy([0], width=w) as strat:
        got = s.copy()
        strat.perm_batch(got)
        assert np.array_equal(got, oracle.perm_batch(s, w))
        one = s[:1].copy()
        strat.perm(one[0])
        assert np.array_equal(one[0], got[0])
        with pytest.raises(HadesError):
            strat.merkle_root(np.zeros((4, 4), dtype=np.uint64))   # compositions are width-5 only
        with pytest.raises(HadesError):
            strat.set_variant(0, 0)
        assert strat.kernel_info("perm")["regs_per_thread"] > 0


_CUSTOM_CONSTANTS_SCRIPT = r"""
import ctypes, sys
import numpy as np
from hades252_b200 import _native
from oracle import cpu_oracle as C

w = int(sys.argv[1])
L, O = _native.lib(), C.lib()
ark = C.gen_elems(4242, 960)                 # random round constants (< 2^254 < p, Montgomery limbs)
mds = C.gen_elems(777 + w, w * w)            # random dense matrix instead of the Cauchy MDS
n = 1500
states = C.gen_elems(99, w * n).reshape(n, w, 4)
want = states.copy()
p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
assert O.oracle_perm_batch(p(want), n, w, p(ark), p(mds), 4) == 0
ctx = _native.ctx_p()
dev = (ctypes.c_int * 1)(0)
rc = L.hades_init(ctypes.byref(ctx), dev, 1, w, p(ark), 960, p(mds))
assert rc == 0, L.hades_last_error(None)
for algo in (2, 1, 0):
    assert L.hades_set_variant(ctx, algo, 0) == 0, L.hades_last_error(ctx)
    got = states.copy()
    assert L.hades_perm_batch(ctx, got.ctypes.data_as(ctypes.c_void_p), n) == 0, L.hades_last_error(ctx)
    assert np.array_equal(got, want), f"algo {algo} differs from the oracle with custom constants"
L.hades_destroy(ctx)
print("custom constants OK")
"""


@pytest.mark.parametrize("w", [3, 5, 9])
def test_custom_constants_all_schedules(w):
    """hades_init with RANDOM round constants and a RANDOM dense matrix (own process: the constant tables are
    process-wide per device, like the crate's consts): the host-side derivations of the sparse and the gauged
    canonical-form schedules are generic, so every schedule must match the oracle run with the same tables."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", _CUSTOM_CONSTANTS_SCRIPT, str(w)], cwd=root, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0 and "custom constants OK" in res.stdout, res.stdout + res.stderr


# ---- ragged Merkle tree + openings (partially filled nodes under a bitmask, paths) -------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 17, 63, 64, 65, 1000, 4097, 4 ** 8 + 3])
def test_merkle_root_ragged(cuda_strategy, oracle, n):
    leaves = oracle.gen_elems(31 + n, n)
    got = cuda_strategy.merkle_root_ragged(leaves)
    want = oracle.merkle_tree(leaves)[-1] if n > 1 else leaves[0]
    assert np.array_equal(got, want)
    if n in (4, 64):
        assert np.array_equal(got, cuda_strategy.merkle_root(leaves))


@pytest.mark.parametrize("algo,regs", [(2, 6), (2, 0), (1, 6), (0, 0)])
def test_merkle_tree_and_openings_device(oracle, H, algo, regs):
    import torch
    from hades252_b200 import CudaStrategy
    n = 3 * 4 ** 6 + 1234  # 13522 leaves: ragged on several levels
    leaves = oracle.gen_elems(777, n)
    want_tree = oracle.merkle_tree(leaves)
    with CudaStrategy([0]) as s:
        s.set_variant(algo, regs)
        assert s.merkle_tree_nodes(n) == want_tree.shape[0]
        d_leaves = torch.from_numpy(leaves.view(np.int64)).cuda()
        d_tree = torch.empty((want_tree.shape[0], 4), dtype=torch.int64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        s.merkle_tree_device(d_leaves.data_ptr(), n, d_tree.data_ptr(), stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_tree.cpu().numpy().view(np.uint64), want_tree)
        # openings of a few hundred leaves, including both ends and the ragged tail
        idx = np.unique(np.concatenate([[0, 1, 3, 4, n - 1, n - 2, n - 5], np.arange(0, n, 53)])).astype(np.uint64)
        levels = len(oracle.merkle_level_sizes(n))
        d_idx = torch.from_numpy(idx.view(np.int64)).cuda()
        d_branch = torch.empty((idx.shape[0], levels, 4, 4), dtype=torch.int64, device="cuda")
        s.merkle_open_device(d_leaves.data_ptr(), d_tree.data_ptr(), n, d_idx.data_ptr(), idx.shape[0], d_branch.data_ptr(), stream)
        torch.cuda.synchronize()
        branch = d_branch.cpu().numpy().view(np.uint64)
        for o, i in enumerate(idx):
            assert np.array_equal(branch[o], oracle.merkle_opening(leaves, want_tree, int(i)))
        # one path re-verified from scratch with the big-int reference (root recomputed from the branch)
        root = H.from_mont_limbs([int(x) for x in want_tree[-1]])
        o = len(idx) // 2
        path = [[H.from_mont_limbs([int(x) for x in node]) for node in group] for group in branch[o]]
        leaf = H.from_mont_limbs([int(x) for x in leaves[int(idx[o])]])
        assert H.merkle_verify(leaf, int(idx[o]), n, path, root)


def test_merkle_tree_single_leaf_and_errors(cuda_strategy, oracle):
    from hades252_b200 import HadesError
    leaf = oracle.gen_elems(5, 1)
    assert np.array_equal(cuda_strategy.merkle_root_ragged(leaf), leaf[0])
    assert cuda_strategy.merkle_tree_nodes(1) == 0
    with pytest.raises(HadesError):
        cuda_strategy.merkle_root_ragged(np.empty((0, 4), dtype=np.uint64))


# ---- cooperative small-batch kernels (coop.cuh): one state per 8 lanes, or per warp -------------------------
@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 63, 64, 65, 1000, 4735, 4736])
def test_coop_perm_bit_identical(oracle, n, wide):
    """Batches at or below the cooperative threshold run perm_batch_coop_kernel<8> (or <32>, a warp per state, below
    the wide threshold); same bits as the oracle and as the one-thread-per-state kernel (threshold 0)."""
    import torch
    from hades252_b200 import CudaStrategy
    s = oracle.gen_elems(1234 + n, 5 * n).reshape(n, 5, 4)
    want = oracle.perm_batch(s)
    with CudaStrategy([0]) as strat:
        assert strat.kernel_info("perm_coop")["local_bytes"] == 0
        assert strat.kernel_info("perm_coop_wide")["local_bytes"] == 0
        strat.set_coop_threshold(1 << 20)
        strat.set_coop_wide_threshold((1 << 20) if wide else 0)
        l0 = strat.launch_count
        d = torch.from_numpy(s.view(np.int64).copy()).cuda()
        strat.perm_batch_device(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert strat.launch_count == l0 + 1
        assert np.array_equal(d.cpu().numpy().view(np.uint64).reshape(n, 5, 4), want)
        got = s.copy()
        strat.perm_batch(got)          # host path, small batch -> one cooperative launch
        assert np.array_equal(got, want)
        strat.set_coop_threshold(0)
        got = s.copy()
        strat.perm_batch(got)
        assert np.array_equal(got, want)


def test_coop_single_perm_and_edge_values(oracle, H):
    """`Strategy::perm` (a batch of one) goes through the cooperative kernel by default; edge operands."""
    from hades252_b200 import CudaStrategy
    P = H.P
    vals = [0, 1, P - 1, P - 2, H.R, (1 << 255) % P, (1 << 64) - 1, P >> 1]
    with CudaStrategy([0]) as strat:
        for wide in (592, 0):  # the default (a warp per state for a lone permutation), then the 8-lane kernel
            strat.set_coop_wide_threshold(wide)
            for k in range(len(vals)):
                st = np.array([H.to_mont_limbs(vals[(k + j) % len(vals)]) for j in range(5)], dtype=np.uint64)
                want = oracle.perm_batch(st[None])[0]
                strat.perm(st)
                assert np.array_equal(st, want)


@pytest.mark.parametrize("n", [4 ** 6, 3 * 4 ** 5 + 77, 13])
def test_coop_merkle_levels(oracle, n):
    """every level of the tree through merkle_level_coop_kernel (threshold above the leaf count), ragged tails
    included; root equals the oracle's and the one-thread kernel's"""
    from hades252_b200 import CudaStrategy
    leaves = oracle.gen_elems(4321 + n, n)
    want = oracle.merkle_tree(leaves)[-1]
    with CudaStrategy([0]) as strat:
        strat.set_coop_threshold(1 << 20)
        for wide in (0, 592, 1 << 20):  # 8 lanes per node everywhere / a warp per node on the top levels / everywhere
            strat.set_coop_wide_threshold(wide)
            assert np.array_equal(strat.merkle_root_ragged(leaves), want)
            if n == 4 ** 6:
                assert np.array_equal(strat.merkle_root(leaves), want)
        strat.set_coop_threshold(0)
        assert np.array_equal(strat.merkle_root_ragged(leaves), want)


# ---- sponge with domain separation --------------------------------------------------------------------------
def test_sponge_domain_separation(cuda_strategy, oracle, golden, H):
    for s in golden["sponge_ds"]:
        msg = [int(x, 16) for x in s["message"]]
        elems = np.array([H.to_mont_limbs(x) for x in msg], dtype=np.uint64).reshape(-1, 4)
        offsets = np.array([0, len(msg)], dtype=np.uint64)
        tag = limbs_to_array([s["domain_mont_limbs"]])[0]
        got = cuda_strategy.sponge_batch(elems, offsets, domain_tag=tag)
        assert [int(x) for x in got[0]] == [int(l, 16) for l in s["digest_mont_limbs"]], (s["message"], s["domain"])
    rng = np.random.default_rng(21)
    lens = rng.integers(0, 33, size=5000)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    elems = oracle.gen_elems(77, int(offsets[-1]))
    tag = oracle.gen_elems(123456, 1)[0]
    got = cuda_strategy.sponge_batch(elems, offsets, domain_tag=tag)
    assert np.array_equal(got, oracle.sponge_batch(elems, offsets, domain_tag=tag))
    assert np.array_equal(cuda_strategy.sponge_batch(elems, offsets, domain_tag=np.zeros(4, np.uint64)),
                          cuda_strategy.sponge_batch(elems, offsets))
    from hades252_b200 import HadesError
    with pytest.raises(HadesError):   # the tag must be a canonical field element
        cuda_strategy.sponge_batch(elems, offsets, domain_tag=np.full(4, 2**64 - 1, dtype=np.uint64))
    with pytest.raises(ValueError):   # CSR that points past the element array
        cuda_strategy.sponge_batch(elems[:10], offsets)


# ---- host paths: pageable memory through the pinned bounce buffers, copy probe, current device ---------------
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_host_paths_pageable_and_pinned(oracle, mode):
    """hades_perm_batch on PAGEABLE numpy memory: automatic (bounce pipeline once the batch exceeds one chunk),
    forced bounce, forced direct -- identical outputs; 700 001 states = several 96 MB chunks with a ragged tail"""
    from hades252_b200 import CudaStrategy
    n = 700001 if mode != 1 else 150001
    s = oracle.gen_elems(9000 + mode, 5 * n).reshape(n, 5, 4)
    want = oracle.perm_batch(s)
    with CudaStrategy([0]) as strat:
        strat.set_host_path(mode)
        got = s.copy()
        strat.perm_batch(got)
        assert np.array_equal(got, want)
        path = strat.last_host_path
        assert ("bounce" in path) == (mode in (0, 1)), path
        # the probe runs the same pipeline without the kernel and leaves the buffer untouched
        before = got.copy()
        strat.copy_probe_ptr(got.ctypes.data, n)
        assert np.array_equal(got, before)


def test_entry_points_restore_current_device(cuda_strategy, oracle):
    import torch
    dev = torch.cuda.current_device()
    s = oracle.gen_elems(3, 5 * 100).reshape(100, 5, 4)
    cuda_strategy.perm_batch(s)
    cuda_strategy.merkle_root(oracle.gen_elems(4, 64))
    assert torch.cuda.current_device() == dev
    if torch.cuda.device_count() > 1:
        from hades252_b200 import CudaStrategy
        torch.cuda.set_device(0)
        with CudaStrategy([1]) as other:
            other.perm_batch(s)
            assert torch.cuda.current_device() == 0


def test_set_variant_rejects_unbuilt_shapes():
    from hades252_b200 import CudaStrategy, HadesError
    with CudaStrategy([0], width=3) as s3:
        with pytest.raises(HadesError):
            s3.set_variant(2, 9)      # 8..10 exist for width 5 only
        s3.set_variant(2, 7)
    with CudaStrategy([0]) as s5:
        with pytest.raises(HadesError):
            s5.set_variant(0, 6)      # the dense schedule has no lockstep build
        s5.set_variant(2, 10)
        assert s5.kernel_info("perm")["regs_per_thread"] > 0
        with pytest.raises(HadesError):
            s5.set_variant(2, 11)


_SINGULAR_CONSTANTS_SCRIPT = r"""
import ctypes, sys
import numpy as np
from hades252_b200 import _native
from oracle import cpu_oracle as C

w = 5
L, O = _native.lib(), C.lib()
ark = C.gen_elems(99, 960)
mds = C.gen_elems(7, w * w)
mds[0:w] = 0                                  # first matrix row zero: no sparse / canonical-form factorisation exists
n = 700
states = C.gen_elems(5, w * n).reshape(n, w, 4)
want = states.copy()
p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
assert O.oracle_perm_batch(p(want), n, w, p(ark), p(mds), 4) == 0
ctx = _native.ctx_p()
dev = (ctypes.c_int * 1)(0)
rc = L.hades_init(ctypes.byref(ctx), dev, 1, w, p(ark), 960, p(mds))
assert rc == 0, L.hades_last_error(None)       # falls back to the dense schedule instead of failing
assert L.hades_set_variant(ctx, 2, 6) == 5 and L.hades_set_variant(ctx, 1, 6) == 5
got = states.copy()
assert L.hades_perm_batch(ctx, got.ctypes.data_as(ctypes.c_void_p), n) == 0, L.hades_last_error(ctx)
assert np.array_equal(got, want)
L.hades_destroy(ctx)
print("singular constants OK")
"""


def test_init_falls_back_to_dense_schedule_for_singular_constants():
    """when the sparse factorisation is singular, hades_init keeps working on the dense schedule"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", _SINGULAR_CONSTANTS_SCRIPT], cwd=root, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "singular constants OK" in res.stdout, res.stdout + res.stderr


# ---- single-process multi-device context: NCCL gather of the subtree roots inside the library -----------------
def test_multi_device_merkle_uses_nccl_allgather(oracle):
    import torch
    from hades252_b200 import CudaStrategy
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    g = 1 << (g.bit_length() - 1)
    with CudaStrategy(list(range(g))) as strat:
        assert strat.collective.startswith("ncclAllGather"), strat.collective
        for depth in (6, 8, 10):
            leaves = oracle.gen_elems(50 + depth, 4 ** depth)
            assert np.array_equal(strat.merkle_root(leaves), oracle.merkle_root(leaves))
        n = 900001
        s = oracle.gen_elems(11, 5 * n).reshape(n, 5, 4)
        got = s.copy()
        strat.perm_batch(got)        # pageable memory, one bounce pipeline thread per device
        assert np.array_equal(got, oracle.perm_batch(s))
    with CudaStrategy([0, 0]) as virt:   # a device listed twice cannot form a communicator: peer copies
        assert virt.collective.startswith("peer copies"), virt.collective
        leaves = oracle.gen_elems(2, 4 ** 7)
        assert np.array_equal(virt.merkle_root(leaves), oracle.merkle_root(leaves))


# ---- BASELINE configs at full size (slow: the CPU oracle needs 20-90 s on 16 threads) --------------------------
def test_config3_merkle_2pow24_leaves_root_equals_oracle(cuda_strategy, oracle):
    """BASELINE configs[2] on one GPU: 4-ary Merkle root over 2^24 synthetic leaves, every level on the device,
    compared with the CPU oracle's root (5 592 405 permutations on the host cores) and with the committed value."""
    import torch
    n = 1 << 24
    d = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    cuda_strategy.gen_elems_device(d.data_ptr(), 0, n, 0x4861646573323532)
    scratch = torch.empty((n // 4 + n // 16 + 8) * 4, dtype=torch.int64, device="cuda")
    out = torch.zeros(4, dtype=torch.int64, device="cuda")
    cuda_strategy.merkle_reduce_device(d.data_ptr(), n, 12, scratch.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint64)
    leaves = d.cpu().numpy().view(np.uint64).reshape(n, 4)
    assert np.array_equal(leaves[:1000], oracle.gen_elems(0, 1000))
    assert [hex(int(x)) for x in got] == ["0x7695019b62c48e7e", "0xa403f682e9373c0", "0xd57e20ff7fb97d67", "0x3eab808a8f6b96a3"]
    assert np.array_equal(got, oracle.merkle_root(leaves))
    # the host entry point (H2D inside) gives the same root
    assert np.array_equal(cuda_strategy.merkle_root(leaves), got)


def test_config4_sponge_2pow22_messages_all_digests(cuda_strategy, oracle):
    """BASELINE configs[3]: 2^22 messages of 1 + (splitmix64(seed2 + i) mod 32) elements; ALL 2^22 digests compared."""
    import torch
    from hades252_b200 import sharding
    n = 1 << 22
    seed2 = 0x4861646573323532 ^ 0x5A5A5A5A
    idx = np.arange(n, dtype=np.uint64)
    z = idx + np.uint64(seed2) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    lens = (z % np.uint64(32)) + np.uint64(1)
    offsets = np.concatenate([[np.uint64(0)], np.cumsum(lens, dtype=np.uint64)])
    total = int(offsets[-1])
    elems = torch.empty(total * 4, dtype=torch.int64, device="cuda")
    cuda_strategy.gen_elems_device(elems.data_ptr(), 0, total, 0x4861646573323532)
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    out = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    cuda_strategy.sponge_batch_device(elems.data_ptr(), d_off.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint64).reshape(n, 4)
    host_elems = elems.cpu().numpy().view(np.uint64).reshape(total, 4)
    want = oracle.sponge_batch(host_elems, offsets)
    assert int(sharding.sponge_perm_counts(offsets).sum()) > 1.9e7
    assert np.array_equal(got, want)


def test_config2_2pow26_states_strided_sample(cuda_strategy, oracle):
    """BASELINE configs[1]: 2^26 states generated and permuted on the device; a fixed strided
    sample of 2^20 states is compared with the oracle, and the digest of all outputs is reproducible."""
    import torch
    n = 1 << 26
    free, _ = torch.cuda.mem_get_info()
    if free < n * 160 + (2 << 30):
        pytest.skip("not enough free device memory for 2^26 states")
    d = torch.empty(n * 20, dtype=torch.int64, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    cuda_strategy.gen_elems_device(d.data_ptr(), 0, n * 5, 0x4861646573323532, sp)
    idx = torch.arange(0, n, n >> 20, device="cuda")
    before = d.view(n, 20)[idx].cpu().numpy().view(np.uint64).reshape(-1, 5, 4)
    cuda_strategy.perm_batch_device(d.data_ptr(), n, sp)
    torch.cuda.synchronize()
    after = d.view(n, 20)[idx].cpu().numpy().view(np.uint64).reshape(-1, 5, 4)
    assert np.array_equal(after, oracle.perm_batch(before))
    dig = torch.zeros(4, dtype=torch.int64, device="cuda")
    cuda_strategy.digest_device(d.data_ptr(), 0, n * 20, dig.data_ptr(), sp)
    torch.cuda.synchronize()
    first = dig.cpu().numpy().copy()
    cuda_strategy.gen_elems_device(d.data_ptr(), 0, n * 5, 0x4861646573323532, sp)
    cuda_strategy.perm_batch_device(d.data_ptr(), n, sp)
    dig.zero_()
    cuda_strategy.digest_device(d.data_ptr(), 0, n * 20, dig.data_ptr(), sp)
    torch.cuda.synchronize()
    assert np.array_equal(first, dig.cpu().numpy())
    del d
    torch.cuda.empty_cache()


# ---- batched verification of openings: open -> verify round trip, tampering, oracle agreement ------------------
@pytest.mark.parametrize("n", [1, 5, 64, 1000, 3 * 4 ** 6 + 1234])
def test_merkle_open_verify_roundtrip_and_tampering(cuda_strategy, oracle, H, n):
    import torch
    leaves = oracle.gen_elems(900 + n, n)
    tree = oracle.merkle_tree(leaves) if n > 1 else np.empty((0, 4), dtype=np.uint64)
    root = tree[-1] if n > 1 else leaves[0]
    levels = len(oracle.merkle_level_sizes(n))
    rng = np.random.default_rng(n)
    idx = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, size=min(n, 400))])).astype(np.uint64)
    k = idx.shape[0]
    sp = torch.cuda.current_stream().cuda_stream
    d_leaves = torch.from_numpy(leaves.view(np.int64).copy()).cuda()
    d_tree = torch.from_numpy(tree.view(np.int64).copy()).cuda() if n > 1 else torch.zeros(4, dtype=torch.int64, device="cuda")
    d_idx = torch.from_numpy(idx.view(np.int64).copy()).cuda()
    d_branch = torch.zeros((k, max(levels, 1), 4, 4), dtype=torch.int64, device="cuda")
    d_root = torch.from_numpy(root.view(np.int64).copy()).cuda()
    d_ok = torch.full((k,), 7, dtype=torch.int32, device="cuda")
    cuda_strategy.merkle_open_device(d_leaves.data_ptr(), d_tree.data_ptr(), n, d_idx.data_ptr(), k, d_branch.data_ptr(), sp)
    cuda_strategy.merkle_verify_device(d_leaves.data_ptr(), n, d_idx.data_ptr(), k, d_branch.data_ptr(), d_root.data_ptr(), d_ok.data_ptr(), sp)
    torch.cuda.synchronize()
    assert d_ok.cpu().numpy().tolist() == [1] * k              # every genuine opening verifies (cooperative kernels: k <= 4736)
    cuda_strategy.set_coop_wide_threshold(0)                   # ... on the 8-lane kernel alone
    d_ok.fill_(7)
    cuda_strategy.merkle_verify_device(d_leaves.data_ptr(), n, d_idx.data_ptr(), k, d_branch.data_ptr(), d_root.data_ptr(), d_ok.data_ptr(), sp)
    torch.cuda.synchronize()
    cuda_strategy.set_coop_wide_threshold(592)
    assert d_ok.cpu().numpy().tolist() == [1] * k
    cuda_strategy.set_coop_threshold(0)                        # ... and on the one-thread kernel
    d_ok.fill_(7)
    cuda_strategy.merkle_verify_device(d_leaves.data_ptr(), n, d_idx.data_ptr(), k, d_branch.data_ptr(), d_root.data_ptr(), d_ok.data_ptr(), sp)
    torch.cuda.synchronize()
    cuda_strategy.set_coop_threshold(4736)
    assert d_ok.cpu().numpy().tolist() == [1] * k
    if levels == 0:
        return
    # tamper: flip one bit of one limb somewhere in half of the branches, a wrong index in some, a wrong root for all
    branch = d_branch.cpu().numpy().view(np.uint64).reshape(k, levels, 4, 4).copy()
    bad = branch.copy()
    victims = rng.choice(k, size=max(1, k // 2), replace=False)
    for v in victims:
        l, c, limb = rng.integers(0, levels), rng.integers(0, 4), rng.integers(0, 4)
        bad[v, l, c, limb] ^= np.uint64(1) << np.uint64(rng.integers(0, 62))
    bad_idx = idx.copy()
    movers = [v for v in rng.choice(k, size=min(k, 20), replace=False) if n > 1]
    for v in movers:
        bad_idx[v] = (int(idx[v]) + 1) % n if n > 1 else idx[v]
    want = oracle.merkle_verify_batch(leaves[bad_idx.astype(np.int64)], bad_idx, n, bad, root)
    assert not want[victims].any()
    d_bad = torch.from_numpy(bad.view(np.int64)).cuda()
    d_bidx = torch.from_numpy(bad_idx.view(np.int64).copy()).cuda()
    cuda_strategy.merkle_verify_device(d_leaves.data_ptr(), n, d_bidx.data_ptr(), k, d_bad.data_ptr(), d_root.data_ptr(), d_ok.data_ptr(), sp)
    torch.cuda.synchronize()
    assert np.array_equal(d_ok.cpu().numpy().astype(bool), want)
    wrong_root = torch.from_numpy(oracle.gen_elems(1, 1)[0].view(np.int64).copy()).cuda()
    cuda_strategy.merkle_verify_device(d_leaves.data_ptr(), n, d_idx.data_ptr(), k, d_branch.data_ptr(), wrong_root.data_ptr(), d_ok.data_ptr(), sp)
    torch.cuda.synchronize()
    assert not d_ok.cpu().numpy().any()
    # one genuine path re-verified from scratch with the big-int reference
    o = k // 2
    path = [[H.from_mont_limbs([int(x) for x in node]) for node in group] for group in branch[o]]
    assert H.merkle_verify(H.from_mont_limbs([int(x) for x in leaves[int(idx[o])]]), int(idx[o]), n, path,
                           H.from_mont_limbs([int(x) for x in root]))


def test_merkle_open_verify_full_size_roundtrip(cuda_strategy, oracle):
    """size-independent property at BASELINE size: resident tree over 2^24 leaves, 2^20 openings gathered on the device,
    every one of them recomputes to the root (12 permutations each); a tampered copy fails everywhere it was touched"""
    import torch
    n, n_open, levels = 1 << 24, 1 << 20, 12
    sp = torch.cuda.current_stream().cuda_stream
    leaves = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    cuda_strategy.gen_elems_device(leaves.data_ptr(), 0, n, 0x4861646573323532, sp)
    tree = torch.empty(cuda_strategy.merkle_tree_nodes(n) * 4, dtype=torch.int64, device="cuda")
    cuda_strategy.merkle_tree_device(leaves.data_ptr(), n, tree.data_ptr(), sp)
    idx = (torch.arange(n_open, dtype=torch.int64, device="cuda") * 2654435761) % n
    branch = torch.empty(n_open * levels * 16, dtype=torch.int64, device="cuda")
    ok = torch.zeros(n_open, dtype=torch.int32, device="cuda")
    cuda_strategy.merkle_open_device(leaves.data_ptr(), tree.data_ptr(), n, idx.data_ptr(), n_open, branch.data_ptr(), sp)
    root = tree[-4:].clone()
    cuda_strategy.merkle_verify_device(leaves.data_ptr(), n, idx.data_ptr(), n_open, branch.data_ptr(), root.data_ptr(), ok.data_ptr(), sp)
    torch.cuda.synchronize()
    assert [hex(int(x)) for x in root.cpu().numpy().view(np.uint64)] == ["0x7695019b62c48e7e", "0xa403f682e9373c0", "0xd57e20ff7fb97d67", "0x3eab808a8f6b96a3"]
    assert int(ok.sum().item()) == n_open
    # flip one bit in the sibling data of every 7th opening (child (pos + 1) % 4 of level 3: never the path node itself)
    b = branch.view(n_open, levels, 4, 4)
    victims = torch.arange(0, n_open, 7, device="cuda")
    pos = ((idx[victims] >> 6) & 3 + 1) % 4
    b[victims, 3, pos, 2] ^= 1 << 17
    cuda_strategy.merkle_verify_device(leaves.data_ptr(), n, idx.data_ptr(), n_open, branch.data_ptr(), root.data_ptr(), ok.data_ptr(), sp)
    torch.cuda.synchronize()
    okc = ok.bool()
    assert not okc[victims].any()
    mask = torch.ones(n_open, dtype=torch.bool, device="cuda")
    mask[victims] = False
    assert okc[mask].all()


def test_coop_sponge_few_messages(oracle):
    """few messages run sponge_coop_kernel (one message per 8 lanes, no bucketing): same digests as the oracle and as the
    bucketed one-thread kernel (threshold 0), with and without a domain tag; a lone hash is one launch"""
    from hades252_b200 import CudaStrategy
    rng = np.random.default_rng(77)
    for n in (1, 3, 4, 5, 33, 700, 4736):
        lens = rng.integers(0, 41, size=n)
        lens[: min(n, 3)] = [0, 4, 40][: min(n, 3)]
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        elems = oracle.gen_elems(50 + n, int(offsets[-1]))
        want = oracle.sponge_batch(elems, offsets)
        tag = oracle.gen_elems(9, 1)[0]
        want_tag = oracle.sponge_batch(elems, offsets, domain_tag=tag)
        with CudaStrategy([0]) as s:
            assert s.kernel_info("sponge_coop")["local_bytes"] == 0
            assert s.kernel_info("sponge_coop_wide")["local_bytes"] == 0
            for wide in (592, 0, 1 << 20):  # default (a warp per message up to 592 messages) / 8 lanes only / a warp always
                s.set_coop_wide_threshold(wide)
                l0 = s.launch_count
                assert np.array_equal(s.sponge_batch(elems, offsets), want)
                assert s.launch_count == l0 + 1
                assert np.array_equal(s.sponge_batch(elems, offsets, domain_tag=tag), want_tag)
            s.set_coop_threshold(0)
            assert np.array_equal(s.sponge_batch(elems, offsets), want)


def test_kernel_choice_never_changes_the_bits(oracle):
    """hypothesis: any batch size around the cooperative threshold, any threshold, device-resident or host call -- the
    launcher picks the cooperative or the one-thread kernel and the outputs are the oracle's either way"""
    import torch
    from hypothesis import given, settings, strategies as st
    from hades252_b200 import CudaStrategy
    pool = oracle.gen_elems(2024, 5 * 6000).reshape(6000, 5, 4)
    want_all = oracle.perm_batch(pool)
    strat = CudaStrategy([0])
    try:
        @settings(max_examples=40, deadline=None)
        @given(n=st.integers(1, 6000), thr=st.sampled_from([0, 1, 7, 8, 9, 4735, 4736, 4737, 6000]),
               wide=st.sampled_from([0, 1, 3, 4, 5, 592, 6000]), first=st.integers(0, 100), host=st.booleans())
        def check(n, thr, wide, first, host):
            n = min(n, 6000 - first)
            strat.set_coop_threshold(thr)
            strat.set_coop_wide_threshold(wide)
            s = pool[first:first + n].copy()
            if host:
                strat.perm_batch(s)
                got = s
            else:
                d = torch.from_numpy(s.view(np.int64)).cuda()
                strat.perm_batch_device(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                got = d.cpu().numpy().view(np.uint64).reshape(n, 5, 4)
            assert np.array_equal(got, want_all[first:first + n])
        check()
    finally:
        strat.close()


def test_sponge_host_path_pipelined_chunks(cuda_strategy, oracle):
    """hades_sponge_batch on a host batch larger than one staging chunk (2^21 elements / 2^19 messages): the range is
    cut into chunks that are staged, uploaded, hashed and read back on rotating streams -- same digests as the oracle,
    incl. empty messages at the chunk edges and one message longer than the others"""
    rng = np.random.default_rng(4)
    n = 330000
    lens = rng.integers(0, 33, size=n)
    lens[[0, 1, n - 1]] = 0
    lens[n // 2] = 5000
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    assert int(offsets[-1]) > 2 * (1 << 21)
    elems = oracle.gen_elems(31, int(offsets[-1]))
    got = cuda_strategy.sponge_batch(elems, offsets)
    assert np.array_equal(got, oracle.sponge_batch(elems, offsets))
    tag = oracle.gen_elems(77, 1)[0]
    got = cuda_strategy.sponge_batch(elems[: int(offsets[200000])], offsets[:200001], domain_tag=tag)
    assert np.array_equal(got, oracle.sponge_batch(elems[: int(offsets[200000])], offsets[:200001], domain_tag=tag))


def test_new_entry_points_reject_bad_arguments(cuda_strategy):
    """argument errors of the round-2 entry points come back as status codes, never as crashes"""
    import ctypes
    import torch
    from hades252_b200 import CudaStrategy, HadesError, _native
    L = _native.lib()
    ctx = cuda_strategy._ctx
    d = torch.zeros(64, dtype=torch.int64, device="cuda")
    ok = torch.zeros(4, dtype=torch.int32, device="cuda")
    # merkle_verify: null pointers, misaligned pointers, zero leaves
    assert L.hades_merkle_verify_dev(ctx, 0, None, 16, d.data_ptr(), 1, d.data_ptr(), d.data_ptr(), ok.data_ptr(), None) == 1
    assert L.hades_merkle_verify_dev(ctx, 0, d.data_ptr(), 0, d.data_ptr(), 1, d.data_ptr(), d.data_ptr(), ok.data_ptr(), None) == 1
    assert L.hades_merkle_verify_dev(ctx, 0, d.data_ptr() + 8, 16, d.data_ptr(), 1, d.data_ptr(), d.data_ptr(), ok.data_ptr(), None) == 1
    assert L.hades_merkle_verify_dev(ctx, 0, d.data_ptr(), 16, None, 1, d.data_ptr(), d.data_ptr(), ok.data_ptr(), None) == 1
    assert L.hades_merkle_verify_dev(ctx, 9, d.data_ptr(), 16, d.data_ptr(), 1, d.data_ptr(), d.data_ptr(), ok.data_ptr(), None) == 1
    assert L.hades_merkle_verify_dev(ctx, 0, d.data_ptr(), 16, d.data_ptr(), 0, d.data_ptr(), d.data_ptr(), ok.data_ptr(), None) == 0  # nothing to do
    # sharded resident root: a single-device context cannot shard
    arr = (ctypes.c_void_p * 1)(d.data_ptr())
    root = np.zeros(4, dtype=np.uint64)
    assert L.hades_merkle_root_sharded_dev(ctx, arr, 12, root.ctypes.data_as(_native.u64p)) == 2      # not a power of 4
    assert L.hades_merkle_root_sharded_dev(ctx, None, 16, root.ctypes.data_as(_native.u64p)) == 1
    # field-op helper: unknown op, null pointers
    iw, ow = ctypes.c_int(), ctypes.c_int()
    assert L.hades_fr_op_shape(99, ctypes.byref(iw), ctypes.byref(ow)) == 1
    assert L.hades_fr_op_dev(ctx, 0, 99, d.data_ptr(), d.data_ptr(), 1, None) == 1
    assert L.hades_fr_op_dev(ctx, 0, 0, None, d.data_ptr(), 1, None) == 1
    # host path knob and cooperative threshold
    assert L.hades_set_host_path(ctx, 3) == 1
    with CudaStrategy([0], width=3) as s3:
        with pytest.raises(HadesError):
            s3.set_coop_threshold(100)      # cooperative kernels exist for width 5 only
    assert L.hades_copy_probe(None, None, 1) == 1
    assert b"" == L.hades_collective(None) and cuda_strategy.collective.startswith("none")
