// Generic-width permutation kernel (runtime W in 2..14): the reference lets a user change WIDTH and the
// MDS asset (README.md:30-31, assets/HOWTO.md) as long as 67*W <= 960 round constants (strategies.rs:40).
// Tuned kernels exist for W = 3, 5, 9 (width_impl.cuh); every other width runs this one: the reference's
// round structure (src/strategies.rs:79-157, src/strategies/scalar.rs:23-49) with the state in a
// runtime-indexed array and the constant tables in global memory.  Correct, bit-identical, not tuned
// (about 5x slower per multiplication than the tuned kernels).
#include <cuda_runtime.h>

#include "fr.cuh"
#include "hades.cuh"
#include "width_ops.hpp"

namespace hades {
namespace {

constexpr int kMaxW = 14;
constexpr int kThreads = 128;

__device__ __forceinline__ void load_fr(Fr& x, const uint32_t* p) {
#pragma unroll
    for (int k = 0; k < 8; k++) x.l[k] = p[k];
}

// tables: ark[67*W][8] then mds[W*W][8] (u32 limbs, Montgomery form), in global memory
__global__ void __launch_bounds__(kThreads) perm_generic_kernel(uint32_t* __restrict__ states, size_t n, int W,
                                                                const uint32_t* __restrict__ tables) {
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const uint32_t* ark = tables;
    const uint32_t* mds = tables + (size_t)kRounds * W * 8;
    uint32_t* p = states + i * (size_t)W * 8;
    Fr s[kMaxW], out[kMaxW];
    for (int j = 0; j < W; j++) load_fr(s[j], p + 8 * j);
    constexpr int kHalf = kFullRounds / 2;
#pragma unroll 1
    for (int r = 0; r < kRounds; r++) {
        for (int j = 0; j < W; j++) {  // scalar.rs:23-30
            Fr c;
            load_fr(c, ark + (size_t)(r * W + j) * 8);
            fr_add(s[j], s[j], c);
        }
        const bool full = r < kHalf || r >= kHalf + kPartialRounds;
        for (int j = full ? 0 : W - 1; j < W; j++) {  // strategies.rs:115 / :89
            Fr x = s[j];
            fr_sbox(x);
            s[j] = x;
        }
        for (int k = 0; k < W; k++) {  // scalar.rs:36-49
            Fr acc;
#pragma unroll
            for (int q = 0; q < 8; q++) acc.l[q] = 0;
#pragma unroll 1
            for (int j = 0; j < W; j++) {
                Fr m, t;
                load_fr(m, mds + (size_t)(k * W + j) * 8);
                fr_mul(t, m, s[j]);
                fr_add(acc, acc, t);
            }
            out[k] = acc;
        }
        for (int j = 0; j < W; j++) s[j] = out[j];
    }
    for (int j = 0; j < W; j++)
#pragma unroll
        for (int k = 0; k < 8; k++) p[8 * j + k] = s[j].l[k];
}

}  // namespace

cudaError_t generic_upload_modulus() { return upload_modulus(); }

cudaError_t generic_launch_perm(uint64_t* d_states, size_t n, int width, const uint64_t* d_tables, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    size_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffULL || width < 2 || width > kMaxW) return cudaErrorInvalidValue;
    perm_generic_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(reinterpret_cast<uint32_t*>(d_states), n, width,
                                                             reinterpret_cast<const uint32_t*>(d_tables));
    return cudaGetLastError();
}

cudaError_t generic_func_attributes(cudaFuncAttributes* out) { return cudaFuncGetAttributes(out, perm_generic_kernel); }

}  // namespace hades
