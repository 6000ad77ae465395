// CUDA kernels of the batched Hades252 engine (sm_100a).  One thread owns one width-W state (or
// one Merkle node / one sponge message) and keeps it in registers across all 67 rounds.
// Memory traffic is 2*32*W bytes per permutation (320 B at W=5) against ~1.5e5 integer multiplies,
// so the kernels are bound by the integer-multiply pipe, not by HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hades.cuh"

namespace hades {

// ---- constant tables (uploaded once per device by hades_init) ------------------------------------
// ROUND_CONSTANTS (src/round_constants.rs:29-48): 960 entries x 8 u32 limbs, Montgomery form.
__constant__ uint32_t c_ark[960 * 8];
// MDS_MATRIX (src/mds_matrix.rs:18-40) for the supported widths, row-major, Montgomery form.
__constant__ uint32_t c_mds3[3 * 3 * 8];
__constant__ uint32_t c_mds5[5 * 5 * 8];
__constant__ uint32_t c_mds9[9 * 9 * 8];

template <int W>
struct DevConsts;
#define HADES_DEVCONSTS(W, SYM)                                                                          \
    template <>                                                                                          \
    struct DevConsts<W> {                                                                                \
        static __device__ __forceinline__ uint32_t ark(int idx, int k) { return c_ark[idx * 8 + k]; }    \
        static __device__ __forceinline__ uint32_t mds(int r, int c, int k) { return SYM[(r * W + c) * 8 + k]; } \
    };
HADES_DEVCONSTS(3, c_mds3)
HADES_DEVCONSTS(5, c_mds5)
HADES_DEVCONSTS(9, c_mds9)

// Montgomery forms of the two small constants the compositions need (checked in tests).
__device__ __forceinline__ void fr_set_one(Fr& x) {  // 1 * R mod p
    const uint32_t v[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                           0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
#pragma unroll
    for (int k = 0; k < 8; k++) x.l[k] = v[k];
}
__device__ __forceinline__ void fr_set_fifteen(Fr& x) {  // 15 * R mod p (Merkle bitmask 0b1111)
    const uint32_t v[8] = {0xffffffdfu, 0x00000020u, 0x00362421u, 0x348ddb9du,
                           0xc2232750u, 0x658b26f6u, 0xa2b2d9b1u, 0x0e5d6e47u};
#pragma unroll
    for (int k = 0; k < 8; k++) x.l[k] = v[k];
}
__device__ __forceinline__ void fr_set_zero(Fr& x) {
#pragma unroll
    for (int k = 0; k < 8; k++) x.l[k] = 0;
}

// 32-byte element <-> registers through two 128-bit accesses (pointers are 16-byte aligned).
__device__ __forceinline__ void fr_load(Fr& x, const uint4* p) {
    uint4 a = p[0], b = p[1];
    x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
    x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
}
__device__ __forceinline__ void fr_store(uint4* p, const Fr& x) {
    p[0] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
    p[1] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
}

constexpr int kPermThreads = 128;

// ---- perm_batch: `Strategy::perm` (strategies.rs:140) over n independent states, in place ---------
template <int W>
__global__ void __launch_bounds__(kPermThreads, 4) perm_batch_kernel(uint4* __restrict__ states, size_t n) {
    size_t i = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    if (i >= n) return;
    uint4* p = states + i * (2 * W);
    Fr s[W];
#pragma unroll
    for (int j = 0; j < W; j++) fr_load(s[j], p + 2 * j);
    hades_perm<W, DevConsts<W>>(s);
#pragma unroll
    for (int j = 0; j < W; j++) fr_store(p + 2 * j, s[j]);
}

// ---- merkle level: out[i] = perm([15, in[4i], in[4i+1], in[4i+2], in[4i+3]])[1] ---------------------
__global__ void __launch_bounds__(kPermThreads, 4)
merkle_level_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_out) {
    size_t i = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    if (i >= n_out) return;
    Fr s[5];
    fr_set_fifteen(s[0]);
    const uint4* p = in + i * 8;  // 4 children = 128 contiguous bytes
#pragma unroll
    for (int j = 0; j < 4; j++) fr_load(s[1 + j], p + 2 * j);
    hades_perm<5, DevConsts<5>>(s);
    fr_store(out + i * 2, s[1]);
}

// ---- sponge: rate 4 / capacity 1, one message per thread (CSR offsets) ------------------------------
// `order` (optional) maps thread -> message so that a warp works on messages of equal block count.
__global__ void __launch_bounds__(kPermThreads, 4)
sponge_kernel(const uint4* __restrict__ elems, const uint64_t* __restrict__ offsets,
              const uint32_t* __restrict__ order, uint4* __restrict__ out, size_t n_msgs) {
    size_t t = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    if (t >= n_msgs) return;
    size_t m = order ? order[t] : t;
    uint64_t b = offsets[m], e = offsets[m + 1];
    Fr s[5];
#pragma unroll
    for (int j = 0; j < 5; j++) fr_set_zero(s[j]);
    bool padded = false;
#pragma unroll 1
    while (!padded) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            Fr x;
            bool add = true;
            if (b < e) {
                fr_load(x, elems + b * 2);
                b++;
            } else if (!padded) {
                fr_set_one(x);
                padded = true;
            } else {
                add = false;
            }
            if (add) fr_add(s[1 + k], s[1 + k], x);
        }
        hades_perm<5, DevConsts<5>>(s);
    }
    fr_store(out + m * 2, s[1]);
}

// ---- synthetic inputs and digests (measurement helpers) ---------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

__global__ void gen_elems_kernel(uint64_t* __restrict__ out, uint64_t first_elem, size_t n_elems, uint64_t seed) {
    size_t n_limbs = n_elems * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_limbs; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = splitmix64(seed + first_elem * 4 + i);
        out[i] = ((i & 3) == 3) ? (v & 0x3fffffffffffffffULL) : v;
    }
}

__global__ void digest_kernel(const uint64_t* __restrict__ limbs, uint64_t first_limb, size_t n_limbs,
                              unsigned long long* __restrict__ digest) {
    uint64_t x = 0, s = 0, x2 = 0, s2 = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_limbs; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = limbs[i];
        uint64_t h = splitmix64(v ^ splitmix64(first_limb + i));
        x ^= h; s += h; x2 ^= v; s2 += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
        x2 ^= __shfl_xor_sync(0xffffffffu, x2, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicXor(digest + 0, (unsigned long long)x);
        atomicAdd(digest + 1, (unsigned long long)s);
        atomicXor(digest + 2, (unsigned long long)x2);
        atomicAdd(digest + 3, (unsigned long long)s2);
    }
}

// ---- integer-multiply roofline microbenchmark --------------------------------------------------------
// Every thread runs `iters` rounds of 8 independent accumulations per round (enough ILP to cover the
// 4-cycle pipe latency with a few warps per scheduler).  Operands come from memory so nothing folds.
constexpr int kPeakIlp = 8;
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_peak_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                        int iters) {
    uint32_t a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    if (VARIANT == 0) {  // IMAD.WIDE.U32 Rd64 = a*b + Rc64, carry-less
        uint64_t acc[kPeakIlp];
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) acc[k] = in[64 + k];
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int k = 0; k < kPeakIlp; k++)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
            }
        }
        uint64_t r = 0;
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) r ^= acc[k];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32);
    } else if (VARIANT == 1) {  // IMAD.WIDE.U32.X: two independent 4-column carry chains per step
        uint32_t e[9], o[9];
#pragma unroll
        for (int k = 0; k < 9; k++) { e[k] = in[64 + k]; o[k] = in[80 + k]; }
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                cmad4(e, a, b, a, b, b);
                cmad4(o, b, a, b, a, a);
                a += 2;
            }
        }
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) r ^= e[k] ^ o[k];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (VARIANT == 2) {  // IMAD (32-bit low multiply-add)
        uint32_t acc[kPeakIlp];
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) acc[k] = in[64 + k];
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int k = 0; k < kPeakIlp; k++) acc[k] = a * acc[k] + b;
            }
        }
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) r ^= acc[k];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {  // mul.lo + mul.hi pair per product (IMAD + IMAD.HI.U32)
        uint32_t lo[kPeakIlp], hi[kPeakIlp];
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) { lo[k] = in[64 + k]; hi[k] = in[72 + k]; }
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int k = 0; k < kPeakIlp; k++) {
                    uint32_t x = lo[k], y = hi[k];
                    lo[k] = x * y + a;
                    hi[k] = __umulhi(x, y) + b;
                }
            }
        }
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) r ^= lo[k] ^ hi[k];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
    }
}

}  // namespace hades
