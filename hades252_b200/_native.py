"""ctypes binding of libhades_b200.so (C ABI: include/hades_cuda.h).  Fails loudly when the library
has not been built: there is no Python/CPU implementation of the permutation in this package."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HADES_B200_LIB: an alternative build of the same library (kernel A/B experiments under tools/)
LIB_PATH = os.environ.get("HADES_B200_LIB") or os.path.join(_HERE, "lib", "libhades_b200.so")

u64p = ctypes.POINTER(ctypes.c_uint64)
ctx_p = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/hades_cuda.h declares
SIGNATURES = {
    "hades_init": (ctypes.c_int, [ctypes.POINTER(ctx_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_uint32,
                                  u64p, ctypes.c_size_t, u64p]),
    "hades_destroy": (None, [ctx_p]),
    "hades_last_error": (ctypes.c_char_p, [ctx_p]),
    "hades_width": (ctypes.c_uint32, [ctx_p]),
    "hades_device_count": (ctypes.c_int, [ctx_p]),
    "hades_perm_batch": (ctypes.c_int, [ctx_p, ctypes.c_void_p, ctypes.c_size_t]),
    "hades_perm_batch_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "hades_merkle_root": (ctypes.c_int, [ctx_p, ctypes.c_void_p, ctypes.c_size_t, u64p]),
    "hades_merkle_root_sharded_dev": (ctypes.c_int, [ctx_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, u64p]),
    "hades_merkle_reduce_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "hades_merkle_tree_nodes": (ctypes.c_size_t, [ctypes.c_size_t]),
    "hades_merkle_tree_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "hades_merkle_open_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "hades_merkle_verify_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "hades_merkle_root_ragged": (ctypes.c_int, [ctx_p, ctypes.c_void_p, ctypes.c_size_t, u64p]),
    "hades_sponge_batch": (ctypes.c_int, [ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "hades_sponge_batch_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_void_p, ctypes.c_void_p]),
    "hades_sponge_batch_ds": (ctypes.c_int, [ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, u64p, ctypes.c_void_p]),
    "hades_sponge_batch_ds_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                 u64p, ctypes.c_void_p, ctypes.c_void_p]),
    "hades_collective": (ctypes.c_char_p, [ctx_p]),
    "hades_copy_probe": (ctypes.c_int, [ctx_p, ctypes.c_void_p, ctypes.c_size_t]),
    "hades_set_host_path": (ctypes.c_int, [ctx_p, ctypes.c_int]),
    "hades_last_host_path": (ctypes.c_char_p, [ctx_p]),
    "hades_host_register": (ctypes.c_int, [ctx_p, ctypes.c_void_p, ctypes.c_size_t]),
    "hades_host_unregister": (ctypes.c_int, [ctx_p, ctypes.c_void_p]),
    "hades_gen_elems_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t,
                                           ctypes.c_uint64, ctypes.c_void_p]),
    "hades_digest_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "hades_imad_peak": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "hades_kernel_info": (ctypes.c_int, [ctx_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "hades_set_variant": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_int]),
    "hades_set_coop_threshold": (ctypes.c_int, [ctx_p, ctypes.c_size_t]),
    "hades_set_coop_wide_threshold": (ctypes.c_int, [ctx_p, ctypes.c_size_t]),
    "hades_fr_op_shape": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "hades_fr_op_dev": (ctypes.c_int, [ctx_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                       ctypes.c_void_p]),
    "hades_launch_count": (ctypes.c_uint64, [ctx_p]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(hades252_b200/build.py).  There is no CPU fallback for the permutation.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
