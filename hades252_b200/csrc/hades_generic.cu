// Permutation kernels for the widths WITHOUT a tuned build: the reference lets a user change WIDTH and the MDS asset
// (README.md:30-31, assets/HOWTO.md) as long as 67*W <= 960 round constants (strategies.rs:40).  Tuned kernels exist
// for W = 3, 5, 9 (width_impl.cuh); every other width in 2..14 runs perm_dense_kernel<W> below: the reference's round
// structure (src/strategies.rs:79-157, src/strategies/scalar.rs:23-49) compiled per width, state in registers, each
// MDS row ONE lazily reduced W-term Montgomery dot product (fr.cuh dot_mont<W>; the reference reduces every product
// and every addition), constants read from a per-context table in global memory with warp-uniform addresses (one
// broadcast transaction per constant; the 64 KB constant bank is per translation unit and is left to the tuned widths).
// Round 1 ran a single runtime-width kernel with one reduction per matrix entry, about 5x slower per multiplication.
#include <cuda_runtime.h>

#include "fr.cuh"
#include "hades.cuh"
#include "width_ops.hpp"

namespace hades {
namespace {

constexpr int kMaxW = 14;
constexpr int kThreads = 128;

// 8 limbs of table entry `e` (32-byte aligned: the table comes from cudaMalloc)
__device__ __forceinline__ void load_entry(uint32_t (&c)[8], const uint4* __restrict__ tab, int e) {
    const uint4 lo = __ldg(tab + 2 * e), hi = __ldg(tab + 2 * e + 1);
    c[0] = lo.x; c[1] = lo.y; c[2] = lo.z; c[3] = lo.w;
    c[4] = hi.x; c[5] = hi.y; c[6] = hi.z; c[7] = hi.w;
}

// tables: ark[67*W] then mds[W*W], 8 u32 Montgomery limbs each
template <int W>
__device__ __forceinline__ void perm_dense(Fr (&s)[W], const uint4* __restrict__ tab) {
    constexpr int kHalf = kFullRounds / 2;
    constexpr int kMds = kRounds * W;
#pragma unroll 1
    for (int r = 0; r < kRounds; r++) {
        // add_round_key (scalar.rs:23-30)
#pragma unroll
        for (int j = 0; j < W; j++) {
            Fr c;
            load_entry(c.l, tab, r * W + j);
            fr_add(s[j], s[j], c);
        }
        // S-box on every word (strategies.rs:115) or on the last one (strategies.rs:89); a real loop over a rotating
        // register file keeps the code small
        if (r < kHalf || r >= kHalf + kPartialRounds) {
#pragma unroll 1
            for (int j = 0; j < W; j++) {
                Fr x = s[0];
                fr_sbox(x);
                rotate_in<W>(s, x);
            }
        } else {
            fr_sbox(s[W - 1]);
        }
        // mul_matrix (scalar.rs:36-49): row k = sum_j M[k][j] s_j, one reduction per row.
        // Bound: < p (1 + 0.4528 W): W <= 6 -> < 4p, W <= 14 -> < 8p.
        Fr out[W];
#pragma unroll
        for (int j = 0; j < W; j++) out[j] = s[j];  // placeholders
#pragma unroll 1
        for (int row = 0; row < W; row++) {
            uint32_t acc[9];
            const int base = kMds + row * W;
            // the state limbs walk the chains, the constants are consumed one limb per step (hades.cuh DotRoles): only W
            // constant limbs are live per step, fetched as broadcast loads
            uint32_t m[W][8];
#pragma unroll
            for (int j = 0; j < W; j++) load_entry(m[j], tab, base + j);
            dot_mont<W>(acc, [&](int j, int k) { return s[j].l[k]; }, [&](int j, int i) { return m[j][i]; });
            Fr res;
            canon<(W <= 6) ? 1 : 2>(res, acc);
            rotate_in<W>(out, res);
        }
#pragma unroll
        for (int j = 0; j < W; j++) s[j] = out[j];
    }
}

template <int W>
__global__ void __launch_bounds__(kThreads) perm_dense_kernel(uint4* __restrict__ states, size_t n, const uint4* __restrict__ tab) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    uint4* p = states + i * (2 * W);
    Fr s[W];
#pragma unroll
    for (int j = 0; j < W; j++) {
        const uint4 lo = p[2 * j], hi = p[2 * j + 1];
        s[j].l[0] = lo.x; s[j].l[1] = lo.y; s[j].l[2] = lo.z; s[j].l[3] = lo.w;
        s[j].l[4] = hi.x; s[j].l[5] = hi.y; s[j].l[6] = hi.z; s[j].l[7] = hi.w;
    }
    perm_dense<W>(s, tab);
#pragma unroll
    for (int j = 0; j < W; j++) {
        p[2 * j] = make_uint4(s[j].l[0], s[j].l[1], s[j].l[2], s[j].l[3]);
        p[2 * j + 1] = make_uint4(s[j].l[4], s[j].l[5], s[j].l[6], s[j].l[7]);
    }
}

#define HADES_GENERIC_WIDTHS(X) X(2) X(4) X(6) X(7) X(8) X(10) X(11) X(12) X(13) X(14)

}  // namespace

cudaError_t generic_upload_modulus() { return upload_modulus(); }

cudaError_t generic_launch_perm(uint64_t* d_states, size_t n, int width, const uint64_t* d_tables, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const size_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffULL || width < 2 || width > kMaxW) return cudaErrorInvalidValue;
    uint4* st = reinterpret_cast<uint4*>(d_states);
    const uint4* tab = reinterpret_cast<const uint4*>(d_tables);
    switch (width) {
#define HADES_CASE(w) case w: perm_dense_kernel<w><<<(unsigned)blocks, kThreads, 0, s>>>(st, n, tab); break;
        HADES_GENERIC_WIDTHS(HADES_CASE)
#undef HADES_CASE
        default: return cudaErrorInvalidValue;  // 3, 5, 9 have tuned kernels (width_impl.cuh)
    }
    return cudaGetLastError();
}

cudaError_t generic_func_attributes(int width, cudaFuncAttributes* out) {
    switch (width) {
#define HADES_CASE(w) case w: return cudaFuncGetAttributes(out, perm_dense_kernel<w>);
        HADES_GENERIC_WIDTHS(HADES_CASE)
#undef HADES_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hades
