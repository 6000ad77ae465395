// This Source Code Form is subject to the terms of the Mozilla Public
// License, v. 2.0.
//
// `CudaStrategy`: the batched device strategy added beside `ScalarStrategy`
// (reference: src/strategies/scalar.rs) behind the crate's own `Strategy` trait
// (reference: src/strategies.rs:31-163).  Feature-gated exactly like `GadgetStrategy`
// (`#[cfg(feature = "cuda")]` in strategies.rs and lib.rs, cf. strategies.rs:20-27, lib.rs:29-31).
//
// States cross the FFI zero-copy: `BlsScalar` is `Scalar([u64; 4])` holding Montgomery limbs and the
// device arithmetic uses the same R = 2^256 form, so `&mut [[BlsScalar; WIDTH]]` is handed over as a
// `*mut u64`.  There is no CPU fallback: construction fails when no CUDA device is usable.
//
// NOT COMPILED in the build container (no Rust toolchain there).

extern crate std;

use super::ffi;
use super::Strategy;
use crate::{mds_matrix::MDS_MATRIX, round_constants::ROUND_CONSTANTS, WIDTH};
use core::ffi::{c_int, CStr};
use dusk_bls12_381::BlsScalar;
use std::string::String;
use std::vec::Vec;

// The layout the FFI relies on (lib.rs keeps the crate `no_std`; these are compile-time checks).
const _: () = assert!(core::mem::size_of::<BlsScalar>() == 32);
const _: () = assert!(core::mem::align_of::<BlsScalar>() == 8);

/// Error of the device engine: status code of `include/hades_cuda.h` plus its message.
#[derive(Debug)]
pub struct CudaError {
    /// `hades_status`
    pub status: i32,
    /// `hades_last_error`
    pub message: String,
}

/// Implements a Hades252 strategy that permutes batches of `[BlsScalar; WIDTH]` states on NVIDIA
/// B200 GPUs.
pub struct CudaStrategy {
    ctx: *mut ffi::hades_ctx,
}

// The context is only ever used through `&mut self`.
unsafe impl Send for CudaStrategy {}

impl CudaStrategy {
    /// Constructs a new `CudaStrategy` over the given CUDA device ordinals (shards batches over
    /// them).  Uploads `ROUND_CONSTANTS` and `MDS_MATRIX` -- the very tables `ScalarStrategy` uses,
    /// already in Montgomery form -- to each device's constant memory once.
    pub fn new(devices: &[i32]) -> Result<Self, CudaError> {
        let devs: Vec<c_int> = devices.iter().map(|&d| d as c_int).collect();
        let mut ctx: *mut ffi::hades_ctx = core::ptr::null_mut();
        let rc = unsafe {
            ffi::hades_init(
                &mut ctx,
                devs.as_ptr(),
                devs.len() as c_int,
                WIDTH as u32,
                ROUND_CONSTANTS.as_ptr() as *const u64,
                ROUND_CONSTANTS.len(),
                MDS_MATRIX.as_ptr() as *const u64,
            )
        };
        if rc != ffi::HADES_OK {
            return Err(Self::error(core::ptr::null(), rc));
        }
        Ok(Self { ctx })
    }

    fn error(ctx: *const ffi::hades_ctx, rc: c_int) -> CudaError {
        let msg = unsafe { CStr::from_ptr(ffi::hades_last_error(ctx)) };
        CudaError { status: rc, message: msg.to_string_lossy().into_owned() }
    }

    /// Applies the permutation to every state of the batch, in place.  Bit-identical to calling
    /// `ScalarStrategy::perm` on each state.
    pub fn perm_batch(&mut self, states: &mut [[BlsScalar; WIDTH]]) -> Result<(), CudaError> {
        let rc = unsafe {
            ffi::hades_perm_batch(self.ctx, states.as_mut_ptr() as *mut u64, states.len())
        };
        if rc != ffi::HADES_OK {
            return Err(Self::error(self.ctx, rc));
        }
        Ok(())
    }

    /// Root of the ragged 4-ary tree over any number of leaves: a node with `k < 4` children hashes
    /// `perm([2^k - 1, c0, .., 0])[1]` (bitmask of the present children in word 0).
    pub fn merkle_root_ragged(&mut self, leaves: &[BlsScalar]) -> Result<BlsScalar, CudaError> {
        let mut root = BlsScalar::zero();
        let rc = unsafe {
            ffi::hades_merkle_root_ragged(
                self.ctx,
                leaves.as_ptr() as *const u64,
                leaves.len(),
                &mut root as *mut BlsScalar as *mut u64,
            )
        };
        if rc != ffi::HADES_OK {
            return Err(Self::error(self.ctx, rc));
        }
        Ok(root)
    }

    /// Root of the 4-ary Merkle tree over `leaves` (`leaves.len()` must be a power of 4);
    /// node = `perm([15, c0, c1, c2, c3])[1]`.
    pub fn merkle_root(&mut self, leaves: &[BlsScalar]) -> Result<BlsScalar, CudaError> {
        let mut root = BlsScalar::zero();
        let rc = unsafe {
            ffi::hades_merkle_root(
                self.ctx,
                leaves.as_ptr() as *const u64,
                leaves.len(),
                &mut root as *mut BlsScalar as *mut u64,
            )
        };
        if rc != ffi::HADES_OK {
            return Err(Self::error(self.ctx, rc));
        }
        Ok(root)
    }

    /// Sponge digests (rate 4, capacity 1) of messages in CSR form: message `m` is
    /// `elems[offsets[m]..offsets[m + 1]]`.
    pub fn sponge_batch(
        &mut self,
        elems: &[BlsScalar],
        offsets: &[u64],
    ) -> Result<Vec<BlsScalar>, CudaError> {
        assert!(!offsets.is_empty() && *offsets.last().unwrap() as usize <= elems.len());
        let n = offsets.len() - 1;
        let mut out = std::vec![BlsScalar::zero(); n];
        let rc = unsafe {
            ffi::hades_sponge_batch(
                self.ctx,
                elems.as_ptr() as *const u64,
                offsets.as_ptr(),
                n,
                out.as_mut_ptr() as *mut u64,
            )
        };
        if rc != ffi::HADES_OK {
            return Err(Self::error(self.ctx, rc));
        }
        Ok(out)
    }
}

impl CudaStrategy {
    /// Sponge digests with domain separation: the capacity word starts as `domain` instead of zero.
    pub fn sponge_batch_with_domain(
        &mut self,
        domain: BlsScalar,
        elems: &[BlsScalar],
        offsets: &[u64],
    ) -> Result<Vec<BlsScalar>, CudaError> {
        assert!(!offsets.is_empty() && offsets[0] == 0 && *offsets.last().unwrap() as usize <= elems.len());
        let n = offsets.len() - 1;
        let mut out = std::vec![BlsScalar::zero(); n];
        let rc = unsafe {
            ffi::hades_sponge_batch_ds(
                self.ctx,
                elems.as_ptr() as *const u64,
                offsets.as_ptr(),
                n,
                &domain as *const BlsScalar as *const u64,
                out.as_mut_ptr() as *mut u64,
            )
        };
        if rc != ffi::HADES_OK {
            return Err(Self::error(self.ctx, rc));
        }
        Ok(out)
    }

    /// How a multi-device context gathers subtree roots ("ncclAllGather (NCCL x.y.z, ...)").
    pub fn collective(&self) -> String {
        unsafe { CStr::from_ptr(ffi::hades_collective(self.ctx)) }.to_string_lossy().into_owned()
    }
}

impl Drop for CudaStrategy {
    fn drop(&mut self) {
        unsafe { ffi::hades_destroy(self.ctx) }
    }
}

/// `Strategy::perm` on one state is a batch of one: the whole permutation is a single kernel, so
/// the per-round primitives of the trait are never called by the device path.  They are still part
/// of the trait contract, so they forward to `ScalarStrategy` (same field operations) for generic
/// callers that drive rounds by hand.
impl Strategy<BlsScalar> for CudaStrategy {
    fn add_round_key<'b, I>(&mut self, constants: &mut I, words: &mut [BlsScalar])
    where
        I: Iterator<Item = &'b BlsScalar>,
    {
        super::ScalarStrategy::new().add_round_key(constants, words)
    }

    fn quintic_s_box(&mut self, value: &mut BlsScalar) {
        super::ScalarStrategy::new().quintic_s_box(value)
    }

    fn mul_matrix<'b, I>(&mut self, constants: &mut I, values: &mut [BlsScalar])
    where
        I: Iterator<Item = &'b BlsScalar>,
    {
        super::ScalarStrategy::new().mul_matrix(constants, values)
    }

    fn perm(&mut self, data: &mut [BlsScalar]) {
        // same programmer-error behaviour as the reference: a length other than WIDTH panics
        let state: &mut [BlsScalar; WIDTH] =
            data.try_into().expect("Hades252 perm needs exactly WIDTH scalars");
        self.perm_batch(core::slice::from_mut(state))
            .expect("Hades252 CUDA permutation failed");
    }
}
