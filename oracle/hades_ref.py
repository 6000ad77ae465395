"""CPU oracle (big-int) for the Hades252 permutation -- TEST INFRASTRUCTURE ONLY.

This file restates, with Python integers, the algorithm of the reference crate
`dusk-hades 0.24.1` (paths below are relative to /root/reference):

  * `Strategy::perm`                 src/strategies.rs:140-157
  * `apply_full_round`               src/strategies.rs:107-119
  * `apply_partial_round`            src/strategies.rs:79-93
  * `ScalarStrategy::add_round_key`  src/strategies/scalar.rs:23-30
  * `ScalarStrategy::quintic_s_box`  src/strategies/scalar.rs:32-34
  * `ScalarStrategy::mul_matrix`     src/strategies/scalar.rs:36-49
  * `ROUND_CONSTANTS` loader         src/round_constants.rs:29-48
  * `MDS_MATRIX` loader              src/mds_matrix.rs:18-40
  * `u64_from_buffer`                src/lib.rs:33-44
  * asset generators                 assets/HOWTO.md:21-48 (ark), :71-108 (mds)

The field arithmetic itself lives in the third-party crate `dusk-bls12_381 = "0.13"`
(Cargo.toml:12; not vendored under /root/reference).  Its published semantics are
restated here: `BlsScalar` is an element of F_p (p below) stored as 4 little-endian
u64 Montgomery limbs `x*R mod p`, `R = 2^256`, always fully reduced;
`BlsScalar::from_raw(v)` interprets `v` as a canonical integer and converts it to
Montgomery form (`Scalar(v) * R2`); `internal_repr()` exposes the Montgomery limbs.

PARITY UNPINNED: the reference holds no known-answer vector for `perm` (its tests
are determinism / self-consistency only: scalar.rs:62-74, round_constants.rs:55-65,
README.md:50-65) and no Rust toolchain exists in this image, so the reference
cannot be executed here.  What pins this oracle instead: (1) the regenerated
assets are byte-identical to the reference's `assets/ark.bin` / `assets/mds.bin`
(sha256 below, checked in tests), (2) F_p arithmetic is exact, so outputs are a
function of the mathematical values only, (3) the loader semantics above
("reading A": file bytes are canonical integers), (4) agreement with the
independent survey-time vectors in SURVEY.md section 8(c) and with the
independent 4x64-limb C restatement in oracle/hades_cpu.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (hades252_b200/) never does.
"""
from __future__ import annotations

import hashlib
from typing import Iterable, List, Sequence

# README.md:35, strategies.rs:14
P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
R = (1 << 256) % P
R2 = (R * R) % P
R_INV = pow(R, -1, P)

# lib.rs:20-27
TOTAL_FULL_ROUNDS = 8
PARTIAL_ROUNDS = 59
WIDTH = 5
N_CONSTANTS = 960  # round_constants.rs:16

ARK_BIN_SHA256 = "78c427449282315729eaa2e39e1937e0aa0b010c4c38bcbb1d57016011880485"
MDS_BIN_SHA256 = {
    5: "131915cbeae1bde75422cce7fcf7feb9223a4dec370a937a2133c1f998ded0e7",
    3: "ddbc07408d9c315ff9337bd6fe1ac76fe5857edda3a85a6a19bd2be9fec4b6a5",
    9: "0ae88716268ef82da312bc96e8be97a5785b83011d56ef7fe44bbe0dfddd07e0",
}


# --------------------------------------------------------------------------- assets
def gen_ark_bin() -> bytes:
    """assets/HOWTO.md:21-48: SHA-512 chain seeded with b"poseidon-for-plonk";
    c_k = from_bytes_wide(h_k) + c_{k-1}, c_{-1} = 1; written via internal_repr()
    (Montgomery limbs, little-endian)."""
    h = b"poseidon-for-plonk"
    prev = 1
    out = bytearray()
    for _ in range(N_CONSTANTS):
        h = hashlib.sha512(h).digest()
        c = (int.from_bytes(h, "little") + prev) % P
        prev = c
        out += ((c * R) % P).to_bytes(32, "little")
    return bytes(out)


def gen_mds_bin(width: int = WIDTH) -> bytes:
    """assets/HOWTO.md:71-108: Cauchy matrix 1/(xs[i]+ys[j]), xs[i]=i, ys[j]=j+WIDTH,
    row-major, written via internal_repr()."""
    out = bytearray()
    for i in range(width):
        for j in range(width):
            v = pow(i + j + width, -1, P)
            out += ((v * R) % P).to_bytes(32, "little")
    return bytes(out)


def load_constants_canonical(blob: bytes) -> List[int]:
    """round_constants.rs:29-48 / mds_matrix.rs:18-40: four LE u64 per entry
    (lib.rs:33-44) -> `BlsScalar::from_raw`, i.e. the bytes are a canonical integer.
    Returns the field VALUES (canonical integers mod p)."""
    assert len(blob) % 32 == 0
    vals = [int.from_bytes(blob[i:i + 32], "little") for i in range(0, len(blob), 32)]
    assert all(v < P for v in vals)
    return vals


def ark_values() -> List[int]:
    return load_constants_canonical(gen_ark_bin())


def mds_values(width: int = WIDTH) -> List[List[int]]:
    flat = load_constants_canonical(gen_mds_bin(width))
    return [flat[i * width:(i + 1) * width] for i in range(width)]


# ------------------------------------------------------------------- representation
def to_mont_limbs(x: int) -> List[int]:
    """canonical value -> the 4 LE u64 limbs a `BlsScalar` holds in memory."""
    m = (x % P) * R % P
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_mont_limbs(limbs: Sequence[int]) -> int:
    """4 LE u64 Montgomery limbs -> canonical value."""
    m = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
    assert m < P, "BlsScalar limbs must be fully reduced"
    return m * R_INV % P


# ------------------------------------------------------------------------ algorithm
_ARK = None
_MDS = {}


def _consts(width: int):
    global _ARK
    if _ARK is None:
        _ARK = ark_values()
    if width not in _MDS:
        _MDS[width] = mds_values(width)
    return _ARK, _MDS[width]


def quintic_s_box(x: int) -> int:
    """scalar.rs:32-34: value.square().square() * value."""
    x2 = x * x % P
    x4 = x2 * x2 % P
    return x4 * x % P


def mul_matrix(mds: Sequence[Sequence[int]], values: Sequence[int]) -> List[int]:
    """scalar.rs:36-49: result[k] += MDS[k][j] * values[j] (j outer, k inner)."""
    w = len(values)
    result = [0] * w
    for j in range(w):
        for k in range(w):
            result[k] = (result[k] + mds[k][j] * values[j]) % P
    return result


def perm(state: Sequence[int], width: int | None = None) -> List[int]:
    """strategies.rs:140-157 on canonical field values.  `width` defaults to
    len(state); the reference's `mul_matrix` panics unless len == WIDTH
    (scalar.rs:48), other widths follow README.md:30-31 / assets/HOWTO.md."""
    w = len(state) if width is None else width
    if len(state) != w:
        raise ValueError("state length must equal the permutation width")
    ark, mds = _consts(w)
    if (TOTAL_FULL_ROUNDS + PARTIAL_ROUNDS) * w > len(ark):
        raise ValueError("Hades252 out of ARK constants")  # strategies.rs:40
    words = [x % P for x in state]
    it = iter(ark)

    def full_round(words):  # strategies.rs:107-119
        words = [(x + next(it)) % P for x in words]
        words = [quintic_s_box(x) for x in words]
        return mul_matrix(mds, words)

    def partial_round(words):  # strategies.rs:79-93
        words = [(x + next(it)) % P for x in words]
        words[-1] = quintic_s_box(words[-1])
        return mul_matrix(mds, words)

    for _ in range(TOTAL_FULL_ROUNDS // 2):
        words = full_round(words)
    for _ in range(PARTIAL_ROUNDS):
        words = partial_round(words)
    for _ in range(TOTAL_FULL_ROUNDS // 2):
        words = full_round(words)
    return words


# --------------------------------------------------------- build-defined compositions
# (no reference implementation: removed from the crate in 0.7.0, CHANGELOG.md:159-162;
#  conventions as laid down in SURVEY.md section 8(c))
MERKLE_ARITY = 4
MERKLE_BITMASK = 15


def merkle_node(children: Sequence[int]) -> int:
    """parent = perm([0b1111, c0, c1, c2, c3])[1]."""
    assert len(children) == MERKLE_ARITY
    return perm([MERKLE_BITMASK, *children])[1]


def merkle_root(leaves: Sequence[int]) -> int:
    n = len(leaves)
    if n < 1 or (n & (n - 1)) or (n.bit_length() - 1) % 2:
        raise ValueError("number of leaves must be a power of 4")
    level = list(leaves)
    while len(level) > 1:
        level = [merkle_node(level[i:i + 4]) for i in range(0, len(level), 4)]
    return level[0]


# Ragged 4-ary tree (any number of leaves) and openings -- the "partial bitmask" convention of the Merkle
# callers: a level of m nodes has ceil(m/4) parents; a parent with k present
# children hashes perm([2^k - 1, c_0 .. c_{k-1}, 0 ...])[1] (word 0 = bitmask of the present children,
# absent children are zero).  For 4^d leaves every mask is 0b1111 and the root equals merkle_root().
def merkle_node_partial(children: Sequence[int]) -> int:
    k = len(children)
    assert 1 <= k <= MERKLE_ARITY
    return perm([(1 << k) - 1, *children, *([0] * (MERKLE_ARITY - k))])[1]


def merkle_levels(leaves: Sequence[int]) -> List[List[int]]:
    """levels[0] = leaves, ..., levels[-1] = [root]."""
    if len(leaves) < 1:
        raise ValueError("at least one leaf")
    levels = [[x % P for x in leaves]]
    while len(levels[-1]) > 1:
        cur = levels[-1]
        levels.append([merkle_node_partial(cur[i:i + 4]) for i in range(0, len(cur), 4)])
    return levels


def merkle_root_ragged(leaves: Sequence[int]) -> int:
    return merkle_levels(leaves)[-1][0]


def merkle_opening(leaves: Sequence[int], index: int) -> List[List[int]]:
    """Authentication path of leaf `index`: per level the 4 children of the path's parent (the path node
    included, absent children as 0)."""
    levels = merkle_levels(leaves)
    branch, i = [], index
    for cur in levels[:-1]:
        g = 4 * (i // 4)
        branch.append([cur[g + c] if g + c < len(cur) else 0 for c in range(4)])
        i //= 4
    return branch


def merkle_verify(leaf: int, index: int, n_leaves: int, branch: Sequence[Sequence[int]], root: int) -> bool:
    node, i, m = leaf % P, index, n_leaves
    for group in branch:
        if group[i % 4] != node:
            return False
        k = min(4, m - 4 * (i // 4))
        if any(group[c] != 0 for c in range(k, 4)):
            return False
        node = merkle_node_partial(list(group[:k]))
        i, m = i // 4, (m + 3) // 4
    return m == 1 and node == root


def sponge(message: Iterable[int], domain: int = 0) -> int:
    """rate 4 / capacity 1: state [domain, 0, 0, 0, 0]; pad with one 1 then zeros to a multiple of 4;
    per block add the 4 elements into words 1..4 and perm; output word 1.
    `domain` (default 0) is the domain-separation tag carried by the capacity word: different tags give independent
    hash functions over the same permutation (the convention of the out-of-tree callers of `perm`, SURVEY.md 8(f)4:
    capacity element = domain / length tag, rate words = message).  Build-defined like the sponge itself."""
    msg = [m % P for m in message] + [1]
    while len(msg) % 4:
        msg.append(0)
    state = [domain % P] + [0] * (WIDTH - 1)
    for b in range(0, len(msg), 4):
        for k in range(4):
            state[1 + k] = (state[1 + k] + msg[b + k]) % P
        state = perm(state)
    return state[1]


# ----------------------------------------------------------------- synthetic inputs
SEED = 0x4861646573323532  # ASCII "Hades252"
_M64 = 0xFFFFFFFFFFFFFFFF


def splitmix64(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def synth_limb(seed: int, idx: int, l: int) -> int:
    """limb l (0..3) of synthetic field element number idx: SURVEY.md 8(d).
    Top limb masked to 62 bits => value < 2^254 < p; used directly as Montgomery limbs."""
    v = splitmix64((seed + idx * 4 + l) & _M64)
    return v & 0x3FFFFFFFFFFFFFFF if l == 3 else v
