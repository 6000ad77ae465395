"""CPU-only: pins the oracle (tests/ may import oracle/).  The reference has no KAT for perm
(SURVEY.md section 4), so the anchors are: byte-identical assets, the survey-time independent
vectors, agreement of the two restatements (Python big-int vs C 4x64 limbs), and the reference's own
self-consistency tests restated."""
import hashlib
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import limbs_to_array
from oracle import cpu_oracle as C
from oracle import hades_ref as H

REF_ASSETS = "/root/reference/assets"

# computed at survey time by an independent big-int script (canonical values)
SURVEY_VECTORS = {
    (1, 1, 1, 1, 1): (0x71a5b8040ed5c21f5900c854f34748e89dfb577514b9bd816e62e1b3e3f039c3,
                      0x4390d7dec01afe00e2f7e5148b8070d99021df24b53d4bffec7d42433e4b8ca2),
    (17,) * 5: (0x4a335a5be470b8c178e7e78dfd8abcedee607c75afbff0491c074bae3415b320,
                0x5e0f4e5bf6fa474cf727ce87dd64e6a4753f60758bb8273e04715a469ab14f91),
    (19,) * 5: (0x3879d4c316e78b027b5ca0640a324a8268a8948fa258dc7deb24a6208ff3262f,
                0x01b08ccf909450c5451a01627cef45995adf52101f213129135b89b132c2dde6),
    (5000,) * 5: (0x246568a8dca8b3c5e44d952f8816bb6a40d6fb81c9df08af255afbc1cd4fe26e,
                  0x3bc1e30a27eba7efb2f1e74b77551be598cd7b7af679d72522c1da682a4c9869),
    (0, 1, 2, 3, 4): (0x4c78fe2e2cdb6e76b43742b08a782a771258f76f57b5ffe586f2391a0363013a,
                      0x02e47cfe251226d450f518946a0abcf1e7f721c0685a4382cab9409aee71ff9a),
    (0,) * 5: (0x4448679e00a28dd381089245efaab4249e99c5825ceec146d8aac63a3c3bbc95,
               0x739c65cc0abbdca8a7ce87edb2363ac0aaf217903c9b1729e8d9682fa82bf971),
    (1, 2, 3): (0x15e0ee5c90ddf504decc720d5eead40b4e53c24578b7b2a70a8c355aa9af0b10,
                0x521bc1cfb26a9ed1c47d8bbe9d7cc21b5480c2a5db017fc90250636426feed0f),
    tuple(range(1, 10)): (0x133837f4318b723ee190064a6c4195dfb3d7f65f69389ea8111a1a0971befd5c,
                          0x01667aa443175a71e3daf93c154f8976a80914e62ea46ade014f8d2eb37ff918),
}


def test_assets_regenerate_to_reference_fingerprints():
    assert hashlib.sha256(H.gen_ark_bin()).hexdigest() == H.ARK_BIN_SHA256
    for w in (3, 5, 9):
        assert hashlib.sha256(H.gen_mds_bin(w)).hexdigest() == H.MDS_BIN_SHA256[w]
    assert H.gen_ark_bin()[:32].hex() == "c571de34b82b3d979eab2c718b8f7c6e555906587b40220bfdc6227cb0df676d"


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference tree only exists in the build container")
def test_assets_byte_identical_to_reference_files():
    assert open(os.path.join(REF_ASSETS, "ark.bin"), "rb").read() == H.gen_ark_bin()
    assert open(os.path.join(REF_ASSETS, "mds.bin"), "rb").read() == H.gen_mds_bin(5)


def test_asset_encoding_is_montgomery_read_as_canonical():
    """file bytes are c*R mod p; the loader (from_raw) takes them as canonical."""
    mds = H.mds_values(5)
    for i in range(5):
        for j in range(5):
            assert mds[i][j] == pow(i + j + 5, -1, H.P) * H.R % H.P
    # reading-B discriminator must NOT be what we compute
    assert H.perm([1] * 5)[0] != 0x5221c7bb3c002df76daf1d97d2eef86392182eee91e0554079095df74aca0c56


def test_python_oracle_matches_survey_vectors():
    for inp, (w0, wlast) in SURVEY_VECTORS.items():
        out = H.perm(list(inp))
        assert out[0] == w0 and out[-1] == wlast, inp
    assert H.to_mont_limbs(H.perm([1] * 5)[0]) == [0x935feb66a5e6cf3c, 0x2409c7dd1a61ab1c,
                                                    0x832c33cbf2dd481f, 0x23338e018f505a2a]


def test_golden_file_matches_python_oracle(golden):
    assert golden["ark_bin_sha256"] == H.ARK_BIN_SHA256
    for c in golden["perm"]:
        vals = [int(x, 16) for x in c["input"]]
        assert [hex(v) for v in H.perm(vals)] == c["output"], c["name"]
        assert [[int(l, 16) for l in w] for w in c["output_mont_limbs"]] == [H.to_mont_limbs(int(x, 16)) for x in c["output"]]
    for m in golden["merkle"]:
        assert hex(H.merkle_root(list(range(m["leaves"])))) == m["root"]
    for s in golden["sponge"]:
        assert hex(H.sponge([int(x, 16) for x in s["message"]])) == s["digest"]


def test_merkle_and_sponge_survey_vectors():
    assert H.merkle_root(list(range(16))) == 0x47a6d5ba3e68f329308fc9e6b126a61da0330ebf52118187e125650252ff9e48
    assert H.merkle_root(list(range(64))) == 0x4d453954f18374288cedeadca0aacd99116c0ab7571cb86abbb6684e85d5021e
    assert H.sponge([]) == 0x54902de3606a3f4114fa32d122fdb4ad022d38cb08fc56d713220afb0283c700
    assert H.sponge([1]) == 0x58cf4ef4b0c4e659bc84238746c98426cfc329c5c0e2a491310a607d9fdc8a19
    assert H.sponge([1, 2, 3, 4, 5]) == 0x64b5fa238772218cf134d0766ead57bc604c136c8bef8b74a1cd91ce0db96171
    assert H.sponge(list(range(1, 10))) == 0x1e19683302dcd295555cdfd98d9d53cf483af95755ae92e063523834c1ab594f
    with pytest.raises(ValueError):
        H.merkle_root(list(range(8)))


def test_c_oracle_matches_golden(golden):
    for c in golden["perm"]:
        inp = limbs_to_array(c["input_mont_limbs"])[None]
        out = C.perm_batch(inp, c["width"])
        assert np.array_equal(out[0], limbs_to_array(c["output_mont_limbs"])), c["name"]
    for m in golden["merkle"]:
        leaves = np.array([H.to_mont_limbs(i) for i in range(m["leaves"])], dtype=np.uint64)
        assert [int(x) for x in C.merkle_root(leaves)] == [int(l, 16) for l in m["root_mont_limbs"]]
    msgs = [[int(x, 16) for x in s["message"]] for s in golden["sponge"]]
    elems = np.array([H.to_mont_limbs(x) for m in msgs for x in m], dtype=np.uint64).reshape(-1, 4)
    offsets = np.cumsum([0] + [len(m) for m in msgs]).astype(np.uint64)
    dig = C.sponge_batch(elems, offsets)
    for k, s in enumerate(golden["sponge"]):
        assert [int(x) for x in dig[k]] == [int(l, 16) for l in s["digest_mont_limbs"]]


def test_c_oracle_matches_python_on_random_states():
    n = 48
    for w in (3, 5, 9):
        s = C.gen_elems(1000 * w, w * n).reshape(n, w, 4)
        o = C.perm_batch(s, w)
        for i in range(0, n, 5):
            vals = [H.from_mont_limbs(s[i, j]) for j in range(w)]
            assert [H.from_mont_limbs(o[i, j]) for j in range(w)] == H.perm(vals)


def test_c_oracle_threads_agree():
    s = C.gen_elems(7, 5 * 333).reshape(333, 5, 4)
    assert np.array_equal(C.perm_batch(s, nthreads=1), C.perm_batch(s, nthreads=5))


def test_synthetic_generator_matches_python():
    e = C.gen_elems(12345, 9)
    for i in range(9):
        for l in range(4):
            assert int(e[i, l]) == H.synth_limb(H.SEED, 12345 + i, l)
    assert all(H.from_mont_limbs(x) < H.P for x in e)  # valid BlsScalar limbs


_fe = st.integers(min_value=0, max_value=H.P - 1)


@settings(max_examples=300, deadline=None)
@given(_fe, _fe)
def test_c_field_ops_vs_bigint(a, b):
    am, bm = np.array(H.to_mont_limbs(a), dtype=np.uint64), np.array(H.to_mont_limbs(b), dtype=np.uint64)
    assert H.from_mont_limbs(C.fr_mul(am, bm)) == a * b % H.P
    assert H.from_mont_limbs(C.fr_add(am, bm)) == (a + b) % H.P


@pytest.mark.parametrize("a,b", [(0, 0), (1, 1), (H.P - 1, H.P - 1), (H.P - 1, 1), (H.R, H.R2), ((1 << 255) % H.P, H.P - 2)])
def test_c_field_ops_edges(a, b):
    am, bm = np.array(H.to_mont_limbs(a), dtype=np.uint64), np.array(H.to_mont_limbs(b), dtype=np.uint64)
    assert H.from_mont_limbs(C.fr_mul(am, bm)) == a * b % H.P
    assert H.from_mont_limbs(C.fr_add(am, bm)) == (a + b) % H.P


# ---- the reference's own tests, restated on the oracle -------------------------------------------
def test_hades_det():
    """src/strategies/scalar.rs:62-74"""
    x, y, z = H.perm([17] * 5), H.perm([17] * 5), H.perm([19] * 5)
    assert x == y and x != z


def test_round_constants():
    """src/round_constants.rs:55-65: non-zero and to_bytes/from_bytes round trip."""
    ark = H.ark_values()
    assert len(ark) == 960 and all(c != 0 for c in ark)
    for c in ark:
        assert int.from_bytes(c.to_bytes(32, "little"), "little") == c and c < H.P


def test_readme_example():
    """README.md:50-65"""
    inp = [1] * H.WIDTH
    out = H.perm(inp)
    assert out != inp and len(out) == len(inp)


def test_wrong_length_rejected():
    """scalar.rs:48 copy_from_slice panics on a length mismatch; out of constants: strategies.rs:40."""
    with pytest.raises(ValueError):
        H.perm([1, 2, 3, 4], width=5)
    with pytest.raises(ValueError):
        H.perm([1] * 15)  # 67*15 > 960


# ---- ragged Merkle tree and openings (build-defined convention) ---------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 16, 17, 21, 64, 67])
def test_ragged_merkle_python_vs_c(n):
    from oracle import cpu_oracle as C
    leaves = C.gen_elems(900 + n, n)
    vals = [H.from_mont_limbs([int(x) for x in leaf]) for leaf in leaves]
    levels = H.merkle_levels(vals)
    tree = C.merkle_tree(leaves)
    assert tree.shape[0] == sum(len(lv) for lv in levels[1:]) == sum(C.merkle_level_sizes(n))
    flat = [v for lv in levels[1:] for v in lv]
    for node, want in zip(tree, flat):
        assert [int(x) for x in node] == H.to_mont_limbs(want)
    if n in (1, 4, 16, 64):  # all masks 0b1111: same root as the power-of-4 tree
        assert H.merkle_root_ragged(vals) == H.merkle_root(vals)
        if n > 1:
            assert np.array_equal(tree[-1], C.merkle_root(leaves))


def test_merkle_openings_verify_and_reject_tampering():
    from oracle import cpu_oracle as C
    n = 23
    leaves = C.gen_elems(4321, n)
    vals = [H.from_mont_limbs([int(x) for x in leaf]) for leaf in leaves]
    root = H.merkle_root_ragged(vals)
    tree = C.merkle_tree(leaves)
    for i in (0, 3, 4, 19, 20, 22):
        branch = H.merkle_opening(vals, i)
        assert H.merkle_verify(vals[i], i, n, branch, root)
        got = C.merkle_opening(leaves, tree, i)
        assert [[H.from_mont_limbs([int(x) for x in node]) for node in group] for group in got] == branch
        assert not H.merkle_verify((vals[i] + 1) % H.P, i, n, branch, root)
        bad = [list(g) for g in branch]
        bad[-1][0] = (bad[-1][0] + 1) % H.P
        assert not H.merkle_verify(vals[i], i, n, bad, root)
    assert not H.merkle_verify(vals[0], 1, n, H.merkle_opening(vals, 0), root)


def test_sponge_domain_separation_golden(golden):
    """sponge with a domain tag in the capacity word: Python big-int restatement vs the C oracle vs the committed
    vectors; a zero tag is the plain sponge and different tags give different digests."""
    for s in golden["sponge_ds"]:
        msg = [int(x, 16) for x in s["message"]]
        tag = int(s["domain"], 16)
        assert hex(H.sponge(msg, tag)) == s["digest"]
        elems = np.array([H.to_mont_limbs(x) for x in msg], dtype=np.uint64).reshape(-1, 4)
        offsets = np.array([0, len(msg)], dtype=np.uint64)
        got = C.sponge_batch(elems, offsets, domain_tag=np.array(H.to_mont_limbs(tag), dtype=np.uint64))
        assert [int(x) for x in got[0]] == [int(l, 16) for l in s["digest_mont_limbs"]]
    assert H.sponge([1, 2, 3], 0) == H.sponge([1, 2, 3])
    assert len({H.sponge([1, 2, 3], t) for t in (0, 1, 2, 15)}) == 4
    zero = np.zeros(4, dtype=np.uint64)
    elems = C.gen_elems(3, 9)
    off = np.array([0, 4, 9], dtype=np.uint64)
    assert np.array_equal(C.sponge_batch(elems, off, domain_tag=zero), C.sponge_batch(elems, off))


def test_merkle_verify_batch_matches_bigint_reference():
    n = 77
    leaves = C.gen_elems(5, n)
    tree = C.merkle_tree(leaves)
    idx = np.array([0, 3, 4, 63, 64, 76], dtype=np.uint64)
    branch = np.stack([C.merkle_opening(leaves, tree, int(i)) for i in idx])
    ok = C.merkle_verify_batch(leaves[idx.astype(np.int64)], idx, n, branch, tree[-1])
    assert ok.all()
    leaf_vals = [H.from_mont_limbs([int(x) for x in l]) for l in leaves]
    root = H.from_mont_limbs([int(x) for x in tree[-1]])
    for o, i in enumerate(idx):
        path = [[H.from_mont_limbs([int(x) for x in node]) for node in group] for group in branch[o]]
        assert H.merkle_verify(leaf_vals[int(i)], int(i), n, path, root)
    bad = branch.copy()
    bad[2, 1, 2, 0] ^= np.uint64(4)
    ok = C.merkle_verify_batch(leaves[idx.astype(np.int64)], idx, n, bad, tree[-1])
    assert ok.tolist() == [True, True, False, True, True, True]
    path = [[H.from_mont_limbs([int(x) for x in node]) % H.P for node in group] for group in bad[2]]
    assert not H.merkle_verify(leaf_vals[int(idx[2])], int(idx[2]), n, path, root)
