"""Host-side mirror of the reference crate's public surface for the `perm` path.

Reference (paths relative to /root/reference):
  * consts `TOTAL_FULL_ROUNDS`, `PARTIAL_ROUNDS`, `WIDTH`          src/lib.rs:20-27
  * `trait Strategy` with `perm(&mut self, data: &mut [T])`, `rounds()`   src/strategies.rs:31,140-162
  * `ScalarStrategy::new()`                                          src/strategies/scalar.rs:12-20
`CudaStrategy` is the added device strategy (BASELINE.json north_star): same `perm` contract
(in-place, length must equal WIDTH) plus the batched `perm_batch(&mut [[BlsScalar; WIDTH]])`.
Everything runs through the C ABI of libhades_b200.so (include/hades_cuda.h); this module holds
no arithmetic and no CPU fallback.

States are numpy uint64 arrays of Montgomery limbs, shape [n, WIDTH, 4]: exactly the bytes of the
reference's `[[BlsScalar; WIDTH]]`.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _native, constants

TOTAL_FULL_ROUNDS = 8   # lib.rs:22
PARTIAL_ROUNDS = 59     # lib.rs:26
WIDTH = 5               # lib.rs:27


class HadesError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"hades status {status}: {message}")
        self.status = status


class Strategy:
    """strategies.rs:31 -- the algorithm interface; implementors provide `perm`."""

    def perm(self, data) -> None:  # strategies.rs:140
        raise NotImplementedError

    @staticmethod
    def rounds() -> int:  # strategies.rs:160-162
        return TOTAL_FULL_ROUNDS + PARTIAL_ROUNDS


def _as_u64(a: np.ndarray, what: str) -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != np.uint64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{what} must be a C-contiguous numpy uint64 array (Montgomery limbs)")
    return a


class CudaStrategy(Strategy):
    """Batched device strategy.  `devices`: CUDA ordinals to shard over (default: device 0)."""

    def __init__(self, devices: Optional[Sequence[int]] = None, width: int = WIDTH):
        self._lib = _native.lib()
        self._ctx = _native.ctx_p()
        self.width = int(width)
        devs = list(devices) if devices is not None else [0]
        arr = (ctypes.c_int * len(devs))(*devs)
        ark = np.ascontiguousarray(constants.round_constants())
        mds = np.ascontiguousarray(constants.mds_matrix(self.width)) if 2 <= self.width <= 14 else np.zeros((1, 4), np.uint64)
        rc = self._lib.hades_init(ctypes.byref(self._ctx), arr, len(devs), self.width,
                                  ark.ctypes.data_as(_native.u64p), ark.shape[0], mds.ctypes.data_as(_native.u64p))
        if rc:
            msg = self._lib.hades_last_error(None).decode()
            self._ctx = None
            raise HadesError(rc, msg)
        self.devices = devs

    @classmethod
    def new(cls, devices: Optional[Sequence[int]] = None) -> "CudaStrategy":  # scalar.rs:17-19 analogue
        return cls(devices)

    # ------------------------------------------------------------------ plumbing
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.hades_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc:
            raise HadesError(rc, self._lib.hades_last_error(self._ctx).decode())

    # ------------------------------------------------------------------ the reference path
    def perm(self, data: np.ndarray) -> None:
        """`Strategy::perm` on ONE state, in place.  Like the reference's `mul_matrix`
        (scalar.rs:48 `copy_from_slice`), a length other than WIDTH is a programmer error."""
        _as_u64(data, "data")
        if data.shape != (self.width, 4):
            raise ValueError(f"perm needs exactly WIDTH={self.width} scalars of 4 limbs, got shape {data.shape}")
        self._check(self._lib.hades_perm_batch(self._ctx, data.ctypes.data, 1))

    def perm_batch(self, states: np.ndarray) -> None:
        """`perm_batch(&mut [[BlsScalar; WIDTH]])`: n states, in place, host memory."""
        _as_u64(states, "states")
        if states.ndim != 3 or states.shape[1:] != (self.width, 4):
            raise ValueError(f"states must have shape [n, {self.width}, 4], got {states.shape}")
        self._check(self._lib.hades_perm_batch(self._ctx, states.ctypes.data, states.shape[0]))

    def perm_batch_ptr(self, host_ptr: int, n: int) -> None:
        """Same on a raw host address (e.g. a pinned torch tensor's data_ptr())."""
        self._check(self._lib.hades_perm_batch(self._ctx, host_ptr, n))

    def perm_batch_device(self, device_ptr: int, n: int, stream: int = 0, dev_index: int = 0) -> None:
        """Device-resident, asynchronous on `stream` (a cudaStream_t value)."""
        self._check(self._lib.hades_perm_batch_dev(self._ctx, dev_index, device_ptr, n, stream))

    # ------------------------------------------------------------------ compositions of perm
    def merkle_root(self, leaves: np.ndarray) -> np.ndarray:
        """4-ary Merkle root; leaves uint64 [4^k, 4]."""
        _as_u64(leaves, "leaves")
        if leaves.ndim != 2 or leaves.shape[1] != 4:
            raise ValueError("leaves must have shape [n, 4]")
        root = np.empty(4, dtype=np.uint64)
        self._check(self._lib.hades_merkle_root(self._ctx, leaves.ctypes.data, leaves.shape[0],
                                                root.ctypes.data_as(_native.u64p)))
        return root

    def merkle_root_sharded_device(self, leaf_ptrs: Sequence[int], n_leaves: int) -> np.ndarray:
        """root of a tree whose leaves are resident: leaf_ptrs[g] = device pointer of the g-th device's leaf range"""
        arr = (ctypes.c_void_p * len(leaf_ptrs))(*leaf_ptrs)
        root = np.empty(4, dtype=np.uint64)
        self._check(self._lib.hades_merkle_root_sharded_dev(self._ctx, arr, n_leaves, root.ctypes.data_as(_native.u64p)))
        return root

    def merkle_reduce_device(self, nodes_ptr: int, n_nodes: int, levels: int, scratch_ptr: int, out_ptr: int,
                             stream: int = 0, dev_index: int = 0) -> None:
        self._check(self._lib.hades_merkle_reduce_dev(self._ctx, dev_index, nodes_ptr, n_nodes, levels, scratch_ptr,
                                                      out_ptr, stream))

    # ragged tree + openings (include/hades_cuda.h: any number of leaves, partial nodes under their bitmask)
    def merkle_tree_nodes(self, n_leaves: int) -> int:
        return int(self._lib.hades_merkle_tree_nodes(n_leaves))

    def merkle_root_ragged(self, leaves: np.ndarray) -> np.ndarray:
        _as_u64(leaves, "leaves")
        if leaves.ndim != 2 or leaves.shape[1] != 4:
            raise ValueError("leaves must have shape [n, 4]")
        root = np.empty(4, dtype=np.uint64)
        self._check(self._lib.hades_merkle_root_ragged(self._ctx, leaves.ctypes.data, leaves.shape[0],
                                                       root.ctypes.data_as(_native.u64p)))
        return root

    def merkle_tree_device(self, leaves_ptr: int, n_leaves: int, tree_ptr: int, stream: int = 0, dev_index: int = 0) -> None:
        """all interior levels (level 1 first, root last) of the ragged tree into device memory at tree_ptr"""
        self._check(self._lib.hades_merkle_tree_dev(self._ctx, dev_index, leaves_ptr, n_leaves, tree_ptr, stream))

    def merkle_open_device(self, leaves_ptr: int, tree_ptr: int, n_leaves: int, index_ptr: int, n_open: int,
                           branch_ptr: int, stream: int = 0, dev_index: int = 0) -> None:
        """authentication paths [n_open, levels, 4, 4] u64 of the leaves listed at index_ptr (u64 positions)"""
        self._check(self._lib.hades_merkle_open_dev(self._ctx, dev_index, leaves_ptr, tree_ptr, n_leaves, index_ptr, n_open,
                                                    branch_ptr, stream))

    def merkle_verify_device(self, leaves_ptr: int, n_leaves: int, index_ptr: int, n_open: int, branch_ptr: int, root_ptr: int,
                             ok_ptr: int, stream: int = 0, dev_index: int = 0) -> None:
        """ok[o] (uint32) = 1 iff opening o (leaf leaves[index[o]], branch[o]) recomputes to the root at root_ptr"""
        self._check(self._lib.hades_merkle_verify_dev(self._ctx, dev_index, leaves_ptr, n_leaves, index_ptr, n_open, branch_ptr,
                                                      root_ptr, ok_ptr, stream))

    def sponge_batch(self, elems: np.ndarray, offsets: np.ndarray, domain_tag: Optional[np.ndarray] = None) -> np.ndarray:
        """Sponge digests of n messages in CSR form; elems uint64 [total, 4], offsets uint64 [n+1].
        `domain_tag` (uint64 [4], one canonical field element): initial capacity word (domain separation)."""
        _as_u64(offsets, "offsets")
        if offsets.ndim != 1 or offsets.shape[0] < 1:
            raise ValueError("offsets must be a 1-d array of n+1 entries")
        n = offsets.shape[0] - 1
        if elems.size:
            _as_u64(elems, "elems")
            if elems.ndim != 2 or elems.shape[1] != 4:
                raise ValueError("elems must have shape [total, 4]")
        # the C side copies elems[offsets[0] .. offsets[n]): a CSR that points past the array would read out of bounds
        if int(offsets[0]) != 0:
            raise ValueError("offsets[0] must be 0")
        if int(offsets[-1]) > (elems.shape[0] if elems.size else 0):
            raise ValueError(f"offsets[-1]={int(offsets[-1])} exceeds the {elems.shape[0] if elems.size else 0} elements given")
        out = np.empty((n, 4), dtype=np.uint64)
        ep = elems.ctypes.data if elems.size else None
        if domain_tag is None:
            self._check(self._lib.hades_sponge_batch(self._ctx, ep, offsets.ctypes.data, n, out.ctypes.data))
        else:
            tag = np.ascontiguousarray(domain_tag, dtype=np.uint64)
            if tag.shape != (4,):
                raise ValueError("domain_tag must be one field element: uint64 [4]")
            self._check(self._lib.hades_sponge_batch_ds(self._ctx, ep, offsets.ctypes.data, n,
                                                        tag.ctypes.data_as(_native.u64p), out.ctypes.data))
        return out

    def sponge_batch_device(self, elems_ptr: int, offsets_ptr: int, n_msgs: int, out_ptr: int, stream: int = 0,
                            dev_index: int = 0) -> None:
        self._check(self._lib.hades_sponge_batch_dev(self._ctx, dev_index, elems_ptr, offsets_ptr, n_msgs, out_ptr, stream))

    # ------------------------------------------------------------------ measurement helpers
    def gen_elems_device(self, out_ptr: int, first_elem: int, n_elems: int, seed: int, stream: int = 0,
                         dev_index: int = 0) -> None:
        self._check(self._lib.hades_gen_elems_dev(self._ctx, dev_index, out_ptr, first_elem, n_elems,
                                                  seed & 0xFFFFFFFFFFFFFFFF, stream))

    def digest_device(self, limbs_ptr: int, first_limb: int, n_limbs: int, digest_ptr: int, stream: int = 0,
                      dev_index: int = 0) -> None:
        self._check(self._lib.hades_digest_dev(self._ctx, dev_index, limbs_ptr, first_limb, n_limbs, digest_ptr, stream))

    def imad_peak(self, variant: int, dev_index: int = 0) -> float:
        v = ctypes.c_double()
        self._check(self._lib.hades_imad_peak(self._ctx, dev_index, variant, ctypes.byref(v)))
        return v.value

    def fr_op(self, op: int, operands: np.ndarray) -> np.ndarray:
        """test-only: device field routine `op` (include/hades_cuda.h) on uint32 operands [n, in_words] -> [n, out_words]"""
        import torch
        iw, ow = ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.hades_fr_op_shape(op, ctypes.byref(iw), ctypes.byref(ow)))
        a = np.ascontiguousarray(operands, dtype=np.uint32)
        if a.ndim != 2 or a.shape[1] != iw.value:
            raise ValueError(f"op {op} takes [n, {iw.value}] uint32 words, got {a.shape}")
        d_in = torch.from_numpy(a.view(np.int32)).cuda()
        d_out = torch.empty((a.shape[0], ow.value), dtype=torch.int32, device="cuda")
        self._check(self._lib.hades_fr_op_dev(self._ctx, 0, op, d_in.data_ptr(), d_out.data_ptr(), a.shape[0],
                                              torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return d_out.cpu().numpy().view(np.uint32)

    def kernel_info(self, kernel: str) -> dict:
        regs, local, thr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.hades_kernel_info(self._ctx, kernel.encode(), ctypes.byref(regs), ctypes.byref(local),
                                                ctypes.byref(thr)))
        return {"regs_per_thread": regs.value, "local_bytes": local.value, "max_threads_per_block": thr.value}

    def set_variant(self, algo: int = 2, regs: int = 6) -> None:
        """Kernel variant (bit-identical results): algo 0 dense / 1 sparse partial rounds / 2 gauged canonical form
        (default); regs = launch shape (include/hades_cuda.h; 6 = lockstep 128-thread blocks, the default at W = 3, 5)."""
        self._check(self._lib.hades_set_variant(self._ctx, algo, regs))

    def set_coop_threshold(self, max_states: int) -> None:
        """Batches / Merkle levels of at most `max_states` states use the cooperative 8-lanes-per-state kernels
        (width 5, algo 2; 0 disables; bit-identical results)."""
        self._check(self._lib.hades_set_coop_threshold(self._ctx, max_states))

    def set_coop_wide_threshold(self, max_states: int) -> None:
        """Batches / Merkle levels of at most `max_states` states (within the threshold above) use a whole warp per
        state (lowest latency; default 592, 0 disables; bit-identical results)."""
        self._check(self._lib.hades_set_coop_wide_threshold(self._ctx, max_states))

    def copy_probe_ptr(self, host_ptr: int, n: int) -> None:
        """perm_batch's host pipeline without the kernel (bare H2D + D2H ceiling)."""
        self._check(self._lib.hades_copy_probe(self._ctx, host_ptr, n))

    def set_host_path(self, mode: int) -> None:
        """0 automatic, 1 always pinned bounce buffers, 2 always direct copies"""
        self._check(self._lib.hades_set_host_path(self._ctx, mode))

    @property
    def last_host_path(self) -> str:
        return self._lib.hades_last_host_path(self._ctx).decode()

    @property
    def collective(self) -> str:
        return self._lib.hades_collective(self._ctx).decode()

    def host_register(self, ptr: int, nbytes: int) -> None:
        self._check(self._lib.hades_host_register(self._ctx, ptr, nbytes))

    def host_unregister(self, ptr: int) -> None:
        self._check(self._lib.hades_host_unregister(self._ctx, ptr))

    @property
    def launch_count(self) -> int:
        return int(self._lib.hades_launch_count(self._ctx))
