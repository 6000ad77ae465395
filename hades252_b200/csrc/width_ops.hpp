// Interface between the engine (C ABI, hades_engine.cu) and the per-width kernel translation units
// (hades_w{3,5,9}_ccf.cu, hades_w{3,5,9}.cu, hades_w{3,5,9}_dense.cu).  Each (width, algorithm) is its own TU
// because every TU owns a separate 64 KB `__constant__` bank and the derived tables alone need up to 57 KB.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hades {

// Kernel variants (runtime-selectable so that bench/tests can A/B them; all bit-identical):
//   algo 0 = dense schedule (reference round structure, one lazily reduced dot product per MDS row)
//   algo 1 = sparse partial rounds (host_tables.hpp, derive_tables)
//   algo 2 = gauged canonical-form schedule (host_tables.hpp, derive_tables_ccf) -- the default
//   regs 0 = __launch_bounds__(128, 4) (<=128 registers), 1 = (128, 3) (<=168), 2 = (128, 2) (<=255),
//   3 = (128, 5) (<=96); 4 / 5 = lockstep blocks of 256 / 512 threads with one barrier per round;
//   6.. = lockstep 128-thread blocks (W=5: 6 -> 5 blocks/SM [default], 9 -> 4; W=3: 6 -> 7 [default], 7 -> 5;
//   W=9: 6 -> 2, 7 -> 3 [default]); 7/8 at W=5 are 384/640-thread experiments
//   (algo 1 and 2 only; the barrier keeps the warps of a block on the same instruction-cache lines)
struct Variant {
    int algo;
    int regs;
    // Batches (and Merkle levels) of at most this many states run the cooperative 8-lanes-per-state kernels
    // (coop.cuh; width 5, algo 2 only); 0 disables them.  hades_set_coop_threshold.
    int coop_max = 0;
    // ... and batches / levels of at most this many states (and <= coop_max) the warp-per-state version of them
    // (perm and Merkle level only); 0 disables it.  hades_set_coop_wide_threshold.
    int coop_wide_max = 0;
};
constexpr int kPermThreads = 128;

// initial capacity word (word 0) of the sponge state: zero, or a domain tag (8 x u32 Montgomery limbs)
struct SpongeTag {
    uint32_t l[8];
};

struct WidthOps {
    int width;
    int algo;          // 0: dense table (67*W + W*W entries); 1: OptLayout<W>::kEntries; 2: CcfLayout<W>::kEntries
    size_t table_u64;  // entries * 4
    cudaError_t (*upload)(const uint64_t* table);  // to the CURRENT device
    cudaError_t (*launch_perm)(Variant v, uint64_t* d_states, size_t n, cudaStream_t s);
    // width 5 only (nullptr otherwise)
    // n_out == ceil(n_in / 4); the last node of a ragged level hashes its present children under the matching bitmask
    cudaError_t (*launch_merkle_level)(Variant v, const uint64_t* d_in, uint64_t* d_out, size_t n_out, size_t n_in, cudaStream_t s);
    cudaError_t (*launch_sponge)(Variant v, const uint64_t* d_elems, const uint64_t* d_offsets, const uint32_t* d_order,
                                 uint64_t* d_out, size_t n_threads, SpongeTag tag, cudaStream_t s);
    cudaError_t (*launch_merkle_verify)(Variant v, const uint64_t* d_leaves, const uint64_t* d_index, size_t n_open, size_t n_leaves,
                                        int levels, const uint64_t* d_branch, const uint64_t* d_root, uint32_t* d_ok, cudaStream_t s);
    cudaError_t (*func_attributes)(const char* kernel, Variant v, cudaFuncAttributes* out);
    bool (*supports)(Variant v);  // is this launch shape (regs) built for this width / schedule?
};

// one translation unit per (width, algo): hades_w{3,5,9}.cu (sparse), _dense.cu, _ccf.cu
const WidthOps* width_ops_3(int algo);
const WidthOps* width_ops_5(int algo);
const WidthOps* width_ops_9(int algo);

// widths without a tuned build (hades_generic.cu): one dense-schedule kernel per width in 2..14, tables in global memory
cudaError_t generic_upload_modulus();  // to the CURRENT device
cudaError_t generic_launch_perm(uint64_t* d_states, size_t n, int width, const uint64_t* d_tables, cudaStream_t s);
cudaError_t generic_func_attributes(int width, cudaFuncAttributes* out);

// test-only field-arithmetic kernels (hades_frtest.cu): op codes and shapes in that file
int fr_test_shape(int op, int* in_words, int* out_words);
cudaError_t fr_test_launch(int op, const uint32_t* d_in, uint32_t* d_out, size_t n, cudaStream_t s);

}  // namespace hades
