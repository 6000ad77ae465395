// This Source Code Form is subject to the terms of the Mozilla Public
// License, v. 2.0.
//
// Raw bindings of libhades_b200.so (C ABI: include/hades_cuda.h).  One `extern` item per symbol the
// `CudaStrategy` needs; nothing else of the library is bound here.  NOT COMPILED in the build
// container (no Rust toolchain there): kept mechanical so that it can be checked against the header
// line by line.

#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct hades_ctx {
    _private: [u8; 0],
}

pub const HADES_OK: c_int = 0;

#[link(name = "hades_b200")]
extern "C" {
    // include/hades_cuda.h: hades_init
    pub fn hades_init(
        out: *mut *mut hades_ctx,
        devices: *const c_int,
        n_dev: c_int,
        width: u32,
        ark_limbs: *const u64,
        n_ark: usize,
        mds_limbs: *const u64,
    ) -> c_int;
    pub fn hades_destroy(ctx: *mut hades_ctx);
    pub fn hades_last_error(ctx: *const hades_ctx) -> *const c_char;
    pub fn hades_perm_batch(ctx: *mut hades_ctx, host_states: *mut u64, n: usize) -> c_int;
    pub fn hades_perm_batch_dev(
        ctx: *mut hades_ctx,
        dev_index: c_int,
        d_states: *mut u64,
        n: usize,
        stream: *mut c_void,
    ) -> c_int;
    pub fn hades_merkle_root(
        ctx: *mut hades_ctx,
        host_leaves: *const u64,
        n_leaves: usize,
        root: *mut u64,
    ) -> c_int;
    pub fn hades_merkle_tree_nodes(n_leaves: usize) -> usize;
    pub fn hades_merkle_tree_dev(
        ctx: *mut hades_ctx,
        dev_index: c_int,
        d_leaves: *const u64,
        n_leaves: usize,
        d_tree: *mut u64,
        stream: *mut c_void,
    ) -> c_int;
    pub fn hades_merkle_open_dev(
        ctx: *mut hades_ctx,
        dev_index: c_int,
        d_leaves: *const u64,
        d_tree: *const u64,
        n_leaves: usize,
        d_index: *const u64,
        n_open: usize,
        d_branch: *mut u64,
        stream: *mut c_void,
    ) -> c_int;
    pub fn hades_merkle_root_ragged(
        ctx: *mut hades_ctx,
        host_leaves: *const u64,
        n_leaves: usize,
        root: *mut u64,
    ) -> c_int;
    pub fn hades_sponge_batch(
        ctx: *mut hades_ctx,
        elems: *const u64,
        offsets: *const u64,
        n_msgs: usize,
        out: *mut u64,
    ) -> c_int;
    pub fn hades_host_register(ctx: *mut hades_ctx, ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn hades_host_unregister(ctx: *mut hades_ctx, ptr: *mut c_void) -> c_int;
    // sponge with a domain tag in the capacity word
    pub fn hades_sponge_batch_ds(
        ctx: *mut hades_ctx,
        elems: *const u64,
        offsets: *const u64,
        n_msgs: usize,
        domain_tag: *const u64,
        out: *mut u64,
    ) -> c_int;
    // leaves resident on the context's devices: d_leaves[g] = device pointer of the g-th range
    pub fn hades_merkle_root_sharded_dev(
        ctx: *mut hades_ctx,
        d_leaves: *const *const u64,
        n_leaves: usize,
        root: *mut u64,
    ) -> c_int;
    pub fn hades_merkle_verify_dev(
        ctx: *mut hades_ctx,
        dev_index: c_int,
        d_leaves: *const u64,
        n_leaves: usize,
        d_index: *const u64,
        n_open: usize,
        d_branch: *const u64,
        d_root: *const u64,
        d_ok: *mut u32,
        stream: *mut c_void,
    ) -> c_int;
    pub fn hades_set_coop_threshold(ctx: *mut hades_ctx, max_states: usize) -> c_int;
    pub fn hades_set_coop_wide_threshold(ctx: *mut hades_ctx, max_states: usize) -> c_int;
    pub fn hades_collective(ctx: *const hades_ctx) -> *const c_char;
}
