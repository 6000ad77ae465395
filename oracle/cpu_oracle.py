"""ctypes front-end of oracle/hades_cpu.c -- TEST INFRASTRUCTURE ONLY (see that file's
header).  Builds libhades_oracle.so on demand with the committed Makefile.
numpy arrays of uint64 limbs in, numpy arrays out."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import hades_ref

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhades_oracle.so")
_lib = None

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hades_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhades_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_perm_batch.argtypes = [_u64p, ctypes.c_size_t, ctypes.c_int, _u64p, _u64p, ctypes.c_int]
        _lib.oracle_perm_batch.restype = ctypes.c_int
        _lib.oracle_perm.argtypes = [_u64p, ctypes.c_int, _u64p, _u64p]
        _lib.oracle_perm.restype = ctypes.c_int
        _lib.oracle_merkle_root.argtypes = [_u64p, ctypes.c_size_t, _u64p, _u64p, _u64p, ctypes.c_int]
        _lib.oracle_merkle_root.restype = ctypes.c_int
        _lib.oracle_merkle_tree_nodes.argtypes = [ctypes.c_size_t]
        _lib.oracle_merkle_tree_nodes.restype = ctypes.c_size_t
        _lib.oracle_merkle_tree.argtypes = [_u64p, ctypes.c_size_t, _u64p, _u64p, _u64p, ctypes.c_int]
        _lib.oracle_merkle_tree.restype = ctypes.c_int
        _lib.oracle_sponge_batch.argtypes = [_u64p, _u64p, ctypes.c_size_t, _u64p, _u64p, _u64p, ctypes.c_int]
        _lib.oracle_sponge_batch.restype = ctypes.c_int
        _lib.oracle_sponge_batch_ds.argtypes = [_u64p, _u64p, ctypes.c_size_t, _u64p, _u64p, _u64p, _u64p, ctypes.c_int]
        _lib.oracle_sponge_batch_ds.restype = ctypes.c_int
        _lib.oracle_load_table.argtypes = [_u8p, ctypes.c_size_t, _u64p]
        _lib.oracle_load_table.restype = None
        _lib.oracle_gen_elems.argtypes = [_u64p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_uint64]
        _lib.oracle_gen_elems.restype = None
        _lib.oracle_digest.argtypes = [_u64p, ctypes.c_uint64, ctypes.c_size_t, _u64p]
        _lib.oracle_digest.restype = None
        for f in (_lib.oracle_fr_mul, _lib.oracle_fr_add):
            f.argtypes = [_u64p, _u64p, _u64p]
            f.restype = None
        _lib.oracle_from_raw.argtypes = [_u64p, _u64p]
        _lib.oracle_from_raw.restype = None
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


_TABLES = {}


def tables(width: int = 5):
    """(ark_limbs[960,4], mds_limbs[w*w,4]) as the reference's const tables hold them
    in memory: file bytes -> from_raw (round_constants.rs:41, mds_matrix.rs:33)."""
    if width not in _TABLES:
        L = lib()
        ark_b = np.frombuffer(hades_ref.gen_ark_bin(), dtype=np.uint8).copy()
        mds_b = np.frombuffer(hades_ref.gen_mds_bin(width), dtype=np.uint8).copy()
        ark = np.empty((960, 4), dtype=np.uint64)
        mds = np.empty((width * width, 4), dtype=np.uint64)
        L.oracle_load_table(ark_b.ctypes.data_as(_u8p), 960, _p(ark))
        L.oracle_load_table(mds_b.ctypes.data_as(_u8p), width * width, _p(mds))
        _TABLES[width] = (ark, mds)
    return _TABLES[width]


def perm_batch(states: np.ndarray, width: int = 5, nthreads: int | None = None) -> np.ndarray:
    """states: uint64 [n, width, 4] Montgomery limbs; returns a permuted copy."""
    out = np.ascontiguousarray(states, dtype=np.uint64).copy()
    assert out.ndim == 3 and out.shape[1] == width and out.shape[2] == 4
    ark, mds = tables(width)
    rc = lib().oracle_perm_batch(_p(out), out.shape[0], width, _p(ark), _p(mds),
                                 nthreads or host_threads())
    if rc:
        raise ValueError("oracle_perm_batch: bad width")
    return out


def merkle_root(leaves: np.ndarray, nthreads: int | None = None) -> np.ndarray:
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
    assert leaves.ndim == 2 and leaves.shape[1] == 4
    ark, mds = tables(5)
    root = np.empty(4, dtype=np.uint64)
    rc = lib().oracle_merkle_root(_p(leaves), leaves.shape[0], _p(ark), _p(mds), _p(root),
                                  nthreads or host_threads())
    if rc:
        raise ValueError("number of leaves must be a power of 4")
    return root


def sponge_batch(elems: np.ndarray, offsets: np.ndarray, nthreads: int | None = None,
                 domain_tag: np.ndarray | None = None) -> np.ndarray:
    """domain_tag: uint64 [4] Montgomery limbs of the capacity word (None = zero, the plain sponge)"""
    elems = np.ascontiguousarray(elems, dtype=np.uint64).reshape(-1, 4)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = offsets.shape[0] - 1
    ark, mds = tables(5)
    out = np.empty((n, 4), dtype=np.uint64)
    if elems.shape[0] == 0:
        elems = np.zeros((1, 4), dtype=np.uint64)
    if domain_tag is None:
        lib().oracle_sponge_batch(_p(elems), _p(offsets), n, _p(ark), _p(mds), _p(out), nthreads or host_threads())
    else:
        tag = np.ascontiguousarray(domain_tag, dtype=np.uint64)
        lib().oracle_sponge_batch_ds(_p(elems), _p(offsets), n, _p(tag), _p(ark), _p(mds), _p(out),
                                     nthreads or host_threads())
    return out


def merkle_level_sizes(n_leaves: int) -> list:
    """sizes of the interior levels of the ragged tree, level 1 first, root last"""
    sizes, m = [], n_leaves
    while m > 1:
        m = (m + 3) // 4
        sizes.append(m)
    return sizes


def merkle_tree(leaves: np.ndarray, nthreads: int | None = None) -> np.ndarray:
    """Ragged 4-ary tree over any number of leaves (hades_ref.merkle_levels): all interior levels,
    level 1 first, root last, uint64 [nodes, 4]."""
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
    ark, mds = tables(5)
    tree = np.empty((lib().oracle_merkle_tree_nodes(leaves.shape[0]), 4), dtype=np.uint64)
    if lib().oracle_merkle_tree(_p(leaves), leaves.shape[0], _p(ark), _p(mds), _p(tree), nthreads or host_threads()):
        raise ValueError("oracle_merkle_tree: need at least one leaf")
    return tree


def merkle_opening(leaves: np.ndarray, tree: np.ndarray, index: int) -> np.ndarray:
    """branch [levels, 4, 4] of leaf `index` taken from a tree built by merkle_tree (absent children zero)."""
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
    sizes = merkle_level_sizes(leaves.shape[0])
    branch = np.zeros((len(sizes), 4, 4), dtype=np.uint64)
    cur, i, off = leaves, index, 0
    for l, m in enumerate(sizes):
        g = 4 * (i // 4)
        k = min(4, cur.shape[0] - g)
        branch[l, :k] = cur[g:g + k]
        cur, off, i = tree[off:off + m], off + m, i // 4
    return branch


def merkle_verify_batch(leaf_nodes: np.ndarray, index: np.ndarray, n_leaves: int, branch: np.ndarray, root: np.ndarray) -> np.ndarray:
    """hades_ref.merkle_verify over a batch: leaf_nodes [n, 4], index [n], branch [n, levels, 4, 4], root [4] -> bool [n].
    Every level is one perm_batch over all openings."""
    n = leaf_nodes.shape[0]
    node = np.ascontiguousarray(leaf_nodes, dtype=np.uint64).copy()
    i = np.asarray(index, dtype=np.uint64).astype(np.int64).copy()
    good = i < n_leaves
    i[~good] = 0
    m = n_leaves
    masks = np.array([hades_ref.to_mont_limbs((1 << k) - 1) for k in range(5)], dtype=np.uint64)
    for l in range(branch.shape[1]):
        group = branch[:, l]                                    # [n, 4, 4]
        k = np.minimum(4, m - 4 * (i // 4))                     # present children per opening
        pos = i % 4
        good &= np.all(group[np.arange(n), pos] == node, axis=1)
        for c in range(4):
            good &= (c < k) | np.all(group[:, c] == 0, axis=1)
        states = np.empty((n, 5, 4), dtype=np.uint64)
        states[:, 0] = masks[k]
        states[:, 1:] = group
        node = perm_batch(states)[:, 1]
        i //= 4
        m = (m + 3) // 4
    return good & (m == 1) & np.all(node == np.asarray(root, dtype=np.uint64)[None, :], axis=1)


def gen_elems(first_elem: int, n_elems: int, seed: int = hades_ref.SEED) -> np.ndarray:
    out = np.empty((n_elems, 4), dtype=np.uint64)
    lib().oracle_gen_elems(_p(out), first_elem, n_elems, seed & 0xFFFFFFFFFFFFFFFF)
    return out


def digest(limbs: np.ndarray, first_limb: int = 0) -> np.ndarray:
    flat = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1)
    d = np.empty(4, dtype=np.uint64)
    lib().oracle_digest(_p(flat), first_limb, flat.shape[0], _p(d))
    return d


def fr_mul(a, b):
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    r = np.empty(4, dtype=np.uint64)
    lib().oracle_fr_mul(_p(a), _p(b), _p(r))
    return r


def fr_add(a, b):
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    r = np.empty(4, dtype=np.uint64)
    lib().oracle_fr_add(_p(a), _p(b), _p(r))
    return r
