/*
 * hades_cuda.h -- C ABI of the B200-native batched Hades252 engine (libhades_b200.so).
 *
 * This is the drop-in boundary for ONE path of the reference crate dusk-hades 0.24.1:
 * `ScalarStrategy::perm` (reference: src/strategies.rs:140-157 + src/strategies/scalar.rs:23-49),
 * batched.  A Rust `CudaStrategy` (see INTEGRATION.md and bindings/rust/) binds exactly these
 * symbols; nothing here depends on torch, Python or C++ types.
 *
 * Data format (identical to the reference's in-memory `BlsScalar`, so no conversion either side):
 *   field element = 4 little-endian uint64_t limbs of the MONTGOMERY form x*2^256 mod p, fully
 *   reduced (< p);  a width-W state = W consecutive elements (W*32 bytes);  a batch = n consecutive
 *   states (array-of-structs), i.e. the memory of `&mut [[BlsScalar; WIDTH]]`.
 *   Inputs >= p are a caller contract violation (a `BlsScalar` can never hold them).
 *
 * All functions return 0 on success or a non-zero hades_status; hades_last_error() gives the
 * message.  There is no CPU fallback: without a usable CUDA device hades_init fails.
 * Every entry point restores the calling thread's current CUDA device before it returns.
 *
 * Threading: one hades_ctx is single-caller (mirrors `&mut self`); distinct contexts may be used
 * concurrently.  The constant tables live in per-device `__constant__` memory and are therefore
 * shared by all contexts of a process on that device (in the reference they are compile-time
 * constants of the crate): re-initialising a (device,width) with DIFFERENT tables is an error.
 */
#ifndef HADES_CUDA_H
#define HADES_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference: src/lib.rs:20-27 */
#define HADES_TOTAL_FULL_ROUNDS 8
#define HADES_PARTIAL_ROUNDS 59
#define HADES_WIDTH 5
#define HADES_N_ROUND_CONSTANTS 960 /* src/round_constants.rs:16 */

typedef enum hades_status {
    HADES_OK = 0,
    HADES_ERR_INVALID_ARG = 1,   /* null pointer, bad width, misaligned device pointer, ... */
    HADES_ERR_NOT_POWER_OF_4 = 2,/* merkle: number of leaves is not 4^k, k >= 0 */
    HADES_ERR_CUDA = 3,          /* a CUDA runtime call failed; see hades_last_error */
    HADES_ERR_NO_DEVICE = 4,     /* no CUDA device / requested ordinal does not exist */
    HADES_ERR_CONSTANTS = 5,     /* tables differ from the ones already resident for this width */
    HADES_ERR_OUT_OF_CONSTANTS = 6 /* 67*width > n_ark: "Hades252 out of ARK constants", strategies.rs:40 */
} hades_status;

typedef struct hades_ctx hades_ctx;

/*
 * Create a context over `n_dev` CUDA devices (ordinals in `devices`; NULL => device 0 only when
 * n_dev == 1, or devices 0..n_dev-1) and upload the constant tables once.
 *   width      permutation width, 2..14 (67*width round constants must fit the 960 of the reference).  Tuned
 *              kernels exist for 3, 5 (the reference's WIDTH, lib.rs:27) and 9; other widths run a generic,
 *              slower kernel with the reference's round structure.
 *   ark_limbs  n_ark*4 u64: the raw limbs of `ROUND_CONSTANTS` (src/round_constants.rs:29-48), i.e.
 *              AFTER `BlsScalar::from_raw`.  Round r uses entries [r*width, r*width+width).
 *   mds_limbs  width*width*4 u64: raw limbs of `MDS_MATRIX` row-major (src/mds_matrix.rs:18-40).
 * Replaces: the compile-time const tables of the crate (they become `__constant__` memory).
 */
int hades_init(hades_ctx** out, const int* devices, int n_dev, uint32_t width,
               const uint64_t* ark_limbs, size_t n_ark, const uint64_t* mds_limbs);
void hades_destroy(hades_ctx* ctx);

/* Message of the last failure on this context (or of the last failed hades_init when ctx == NULL).
 * Valid until the next call on the same context / thread. */
const char* hades_last_error(const hades_ctx* ctx);

uint32_t hades_width(const hades_ctx* ctx);
int hades_device_count(const hades_ctx* ctx);

/*
 * `CudaStrategy::perm_batch(&mut [[BlsScalar; WIDTH]])`: permute n states in place.
 * HOST pointer.  Sharded in contiguous ranges over the context's devices, each range streamed in
 * chunks (H2D / kernel / D2H overlapped on separate streams).  Synchronous: on return the host
 * buffer holds the outputs.  Pinned (page-locked) host memory makes the copies asynchronous; see
 * hades_host_register.  A batch of at most 16 KB on a single-device context (a lone `Strategy::perm` is
 * 160 bytes) skips the copy engines: it is staged in a mapped page-locked buffer the kernel works on in place.
 * Replaces: a loop of `ScalarStrategy::perm` (strategies.rs:140).
 */
int hades_perm_batch(hades_ctx* ctx, uint64_t* host_states, size_t n);

/*
 * Device-resident variant: d_states is device memory on devices[dev_index], 16-byte aligned.
 * Asynchronous on `stream` (a cudaStream_t, NULL = the legacy default stream); no host sync.
 */
int hades_perm_batch_dev(hades_ctx* ctx, int dev_index, uint64_t* d_states, size_t n, void* stream);

/*
 * 4-ary Merkle root: node = perm([15, c0, c1, c2, c3])[1] (bitmask of present children in word 0,
 * output word 1); leaves are field elements used as level-0 nodes; n_leaves = 4^k.  Width-5 contexts
 * only.  HOST pointers.  With a power-of-two number of devices (and at least 1024 leaves each) the leaf
 * ranges are sharded over the context's devices: every device reduces its range to 1-2 subtree roots, the
 * roots are all-gathered with ncclAllGather (communicator created by hades_init with ncclCommInitAll;
 * see hades_collective) and every device finishes the top levels.  (Build-defined composition: the
 * reference removed its Merkle code in 0.7.0, CHANGELOG.md:159-162.)
 */
int hades_merkle_root(hades_ctx* ctx, const uint64_t* host_leaves, size_t n_leaves, uint64_t root[4]);
/* Same with the leaves already RESIDENT: d_leaves[g] points at leaves [g n/G, (g+1) n/G) in the memory of the
 * context's g-th device (G = hades_device_count, a power of two; >= 1024 leaves per device).  Synchronous. */
int hades_merkle_root_sharded_dev(hades_ctx* ctx, const uint64_t* const* d_leaves, size_t n_leaves, uint64_t root[4]);

/*
 * Device-resident Merkle reduction: hashes `levels` levels, n_nodes -> n_nodes / 4^levels nodes.
 * d_nodes is read only; d_scratch must hold n_nodes/4 + n_nodes/16 elements (32 B each); the
 * result is written to d_out (n_nodes / 4^levels elements).  Asynchronous on `stream`.
 * This is the per-GPU step of the sharded tree: each rank reduces its leaf range to subtree roots,
 * roots are all-gathered (NCCL), and the top levels are reduced with the same call.
 */
int hades_merkle_reduce_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_nodes, size_t n_nodes,
                            int levels, uint64_t* d_scratch, uint64_t* d_out, void* stream);

/*
 * Ragged 4-ary Merkle tree over ANY number of leaves, kept resident for openings (the callers of `perm`
 * in dusk-poseidon hash partially filled nodes under a bitmask of the present children and hand out
 * authentication paths; SURVEY.md 8(f)4 -- build-defined like hades_merkle_root): a level of m nodes has
 * ceil(m/4) parents; a parent with k present children = perm([2^k - 1, c_0 .. c_{k-1}, 0 ...])[1].  For
 * n_leaves = 4^d the root equals hades_merkle_root's.
 *   hades_merkle_tree_nodes(n)  number of interior nodes (all levels above the leaves); 0 for n <= 1
 *   hades_merkle_tree_dev       writes the interior levels to d_tree (level 1 first, root last:
 *                               hades_merkle_tree_nodes(n) elements of 32 B); asynchronous on `stream`
 *   hades_merkle_open_dev       authentication paths of n_open leaves (d_index: their positions) from a
 *                               resident tree: d_branch[o][l][c] (32 B each, l < number of levels, c < 4) =
 *                               child c of the level-(l+1) ancestor of leaf d_index[o], the path node
 *                               included, zero where the child does not exist; asynchronous on `stream`
 *   hades_merkle_verify_dev     recomputes the root from each opening (levels permutations per opening)
 *   hades_merkle_root_ragged    HOST leaves -> root, on the context's first device
 */
size_t hades_merkle_tree_nodes(size_t n_leaves);
int hades_merkle_tree_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_leaves, size_t n_leaves, uint64_t* d_tree,
                          void* stream);
int hades_merkle_open_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_leaves, const uint64_t* d_tree, size_t n_leaves,
                          const uint64_t* d_index, size_t n_open, uint64_t* d_branch, void* stream);
int hades_merkle_root_ragged(hades_ctx* ctx, const uint64_t* host_leaves, size_t n_leaves, uint64_t root[4]);
/*
 * Batched verification of openings (the consumer of hades_merkle_open_dev): for each of n_open openings, walk the
 * path from leaf d_leaves[d_index[o]] up through d_branch[o][l][0..3] (same layout as hades_merkle_open_dev writes):
 * the running node must sit at its position in every group, absent children of a ragged level must be zero, and the
 * node after the last level must equal d_root (one element).  d_ok[o] = 1 / 0.  One thread per opening, `levels`
 * permutations each; asynchronous on `stream`.
 */
int hades_merkle_verify_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_leaves, size_t n_leaves, const uint64_t* d_index,
                            size_t n_open, const uint64_t* d_branch, const uint64_t* d_root, uint32_t* d_ok, void* stream);

/*
 * Sponge hash (rate 4, capacity 1) of n_msgs variable-length messages given in CSR form:
 * message m = elems[offsets[m] .. offsets[m+1]) (field elements, 32 B each).  State [0;5]; the
 * message is padded with a single 1 then zeros to a multiple of 4 (always at least the 1); each
 * block is added into words 1..4 and permuted; digest = word 1.  out: n_msgs*4 u64.  HOST pointers.
 */
int hades_sponge_batch(hades_ctx* ctx, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs,
                       uint64_t* out);
/* Device-resident variant (all pointers on devices[dev_index]); asynchronous on `stream`. */
int hades_sponge_batch_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_elems, const uint64_t* d_offsets,
                           size_t n_msgs, uint64_t* d_out, void* stream);

/*
 * Sponge with domain separation: as hades_sponge_batch, but the capacity word (word 0) starts as
 * `domain_tag` (one canonical field element, Montgomery limbs) instead of zero, so that different tags give
 * independent hash functions over the same permutation (the convention of the callers of `perm`: the capacity
 * element carries the domain / length tag, the rate words the message).  A zero tag equals hades_sponge_batch.
 */
int hades_sponge_batch_ds(hades_ctx* ctx, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs,
                          const uint64_t domain_tag[4], uint64_t* out);
int hades_sponge_batch_ds_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_elems, const uint64_t* d_offsets,
                              size_t n_msgs, const uint64_t domain_tag[4], uint64_t* d_out, void* stream);

/* The collective multi-device contexts gather subtree roots with: "ncclAllGather (NCCL x.y.z, ...)", or
 * "peer copies (...reason...)" when no communicator could be created, or "none (single device)". */
const char* hades_collective(const hades_ctx* ctx);

/* Page-lock / unlock a caller-owned host range so hades_perm_batch copies run asynchronously.  Not required:
 * PAGEABLE buffers larger than one chunk are staged through the context's own pinned buffers by a pool of
 * host threads (HADES_COPY_THREADS, default min(16, cores)). */
int hades_host_register(hades_ctx* ctx, void* ptr, size_t bytes);
int hades_host_unregister(hades_ctx* ctx, void* ptr);

/* ---- measurement helpers (bench / tests); not part of the reference-facing surface ---------- */

/* hades_perm_batch's host pipeline WITHOUT the kernel (same chunking, same buffers, same streams): the bare
 * H2D + D2H copy ceiling of the box, for bench.py's e2e analysis.  The buffer comes back unchanged. */
int hades_copy_probe(hades_ctx* ctx, uint64_t* host_states, size_t n);
/* Force the host path of hades_perm_batch: 0 = automatic (page-locked memory: direct asynchronous copies;
 * pageable: pinned bounce buffers), 1 = always bounce, 2 = always direct.  For A/B measurements. */
int hades_set_host_path(hades_ctx* ctx, int mode);
/* Which host path the last hades_perm_batch / hades_copy_probe took (static string). */
const char* hades_last_host_path(const hades_ctx* ctx);

/* Fill d_out with n_elems synthetic field elements: element e (global index first_elem + i), limb l
 * = splitmix64(seed + 4*e + l), top limb masked to 62 bits (< 2^254 < p).  SURVEY.md 8(d). */
int hades_gen_elems_dev(hades_ctx* ctx, int dev_index, uint64_t* d_out, uint64_t first_elem, size_t n_elems,
                        uint64_t seed, void* stream);
/* Accumulate a 4-word digest of n_limbs u64 into d_digest[4] (device memory, caller-zeroed):
 * [0] ^= H, [1] += H, [2] ^= limb, [3] += limb, H = splitmix64(limb ^ splitmix64(first_limb + i)). */
int hades_digest_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_limbs, uint64_t first_limb, size_t n_limbs,
                     uint64_t* d_digest, void* stream);
/* Integer-multiply roofline microbenchmark on devices[dev_index]: 32x32->64 multiply-accumulates per
 * second.  variant 0: IMAD.WIDE.U32.X carry chains (the production idiom); 1: IMAD.WIDE.U32 with
 * carry-out only; 2: IMAD 32-bit low half only (context: not a full product); 3: IMAD + IMAD.HI.U32
 * pair per product; 4: 16-link carry chains (one landing add per 16 products: the bare pipe rate).
 * Instruction forms validated in tools/microbench.cu. */
int hades_imad_peak(hades_ctx* ctx, int dev_index, int variant, double* products_per_s);
/* Registers per thread / local (spill) bytes / max threads per block of the context's current kernel
 * variant: kernel = "perm" | "merkle" | "sponge" (the last two for width 5), and for the cooperative small-batch
 * kernels of width 5 "perm_coop" | "merkle_coop" | "sponge_coop" | "perm_coop_wide" | "merkle_coop_wide". */
int hades_kernel_info(hades_ctx* ctx, const char* kernel, int* regs_per_thread, int* local_bytes,
                      int* max_threads_per_block);
/* Select the kernel variant (all bit-identical): algo 0 = dense schedule (the reference's round
 * structure), 1 = sparse partial rounds, 2 = gauged canonical-form schedule (default; HADES_ERR_CONSTANTS if it
 * could not be derived for the context's constants); launch shape `regs`: 0/1/2/3 = 128-thread blocks
 * with at most 128/168/255/96 registers per thread, 4/5 = lockstep blocks of 256/512 threads, 6.. = lockstep
 * 128-thread blocks (the default; see hades252_b200/csrc/width_ops.hpp). */
int hades_set_variant(hades_ctx* ctx, int algo, int regs);  /* HADES_ERR_INVALID_ARG for a shape not built for the width */
/* Small-batch path: batches, Merkle levels and sponge calls of at most `max_states` states / messages run the cooperative kernels --
 * one state per group of 8 lanes, 2.4x lower latency (111 us) than the one-thread-per-state kernel, which is
 * latency-bound below ~2^14 states (a lone `Strategy::perm`, strategies.rs:140, is a batch of one).  Width 5,
 * algo 2 only; 0 disables; default 4736 (two 16-state blocks per SM).  Results are bit-identical either way. */
int hades_set_coop_threshold(hades_ctx* ctx, size_t max_states);
/* Batches, Merkle levels, sponge calls and opening verifications of at most `max_states` states / messages / openings
 * (and within the threshold above) give each state a whole
 * warp instead of 8 lanes: all matrix products of a full round and all dot products of a partial round fit one slot
 * each.  Lowest latency for a lone permutation, a quarter of the capacity per block; default 592 (one block of four
 * states per SM), 0 disables.  Bit-identical results. */
int hades_set_coop_wide_threshold(hades_ctx* ctx, size_t max_states);
/* Test-only: evaluate ONE device field routine of fr.cuh on n caller-supplied operand tuples (u32 limbs, device
 * memory): op 0 fr_mul, 1 fr_add, 2 fr_sbox, 3 sqr_mont (raw 9 limbs), 4 mul_wide, 5 redc16, 6 dot_mont<4>,
 * 7 dot_mont_plus<4>, 8 dot_mont<5>, 9 dot_mont<1>, 10..14 canon<0..4>, 15 mul_const_short<4>, 16 fr_mul_lazy.
 * hades_fr_op_shape gives the u32 words per element on each side.  Drives tests/test_gpu_fr.py. */
int hades_fr_op_shape(int op, int* in_words, int* out_words);
int hades_fr_op_dev(hades_ctx* ctx, int dev_index, int op, const uint32_t* d_in, uint32_t* d_out, size_t n, void* stream);
/* Number of kernel launches issued through this context since creation (bench's gpu_launches). */
uint64_t hades_launch_count(const hades_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HADES_CUDA_H */
