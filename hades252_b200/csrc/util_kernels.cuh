// Measurement helpers: synthetic inputs, digests and the integer-multiply roofline microbenchmark.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fr.cuh"

namespace hades {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// element e (global index), limb l = splitmix64(seed + 4e + l), top limb masked to 62 bits
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

// ---- Merkle openings: gather the authentication paths of n_open leaves out of a resident tree -----------
// tree = interior levels of the ragged 4-ary tree, level 1 first, root last (hades_merkle_tree_dev).
// branch[o][l][c] = child c of the level-(l+1) ancestor of leaf index[o] (the path node included), zero when the
// child does not exist.  One thread per 16-byte chunk: HBM-bound gather, coalesced on the output side.
__global__ void merkle_open_kernel(const uint4* __restrict__ leaves, const uint4* __restrict__ tree, size_t n_leaves,
                                   const uint64_t* __restrict__ index, size_t n_open, int levels, uint4* __restrict__ branch) {
    const size_t total = n_open * (size_t)levels * 8;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int half = (int)(t & 1), c = (int)((t >> 1) & 3);
        const size_t ol = t >> 3;
        const int l = (int)(ol % levels);
        const size_t o = ol / levels;
        // size of level l and offset of level l inside `tree` (level 0 = the leaves)
        size_t m = n_leaves, off = 0;
        for (int k = 0; k < l; k++) {
            m = (m + 3) / 4;
            if (k + 1 < l) off += m;
        }
        // after the loop: m = size of level l; off = sum of sizes of levels 1 .. l-1
        const size_t g = 4 * ((index[o] >> (2 * l)) / 4) + c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (g < m) v = (l == 0 ? leaves : tree + off * 2)[g * 2 + half];
        branch[t] = v;
    }
}

// sponge length bucketing: key = permutations needed by message m = floor(len / 4) + 1, value = m
__global__ void sponge_keys_kernel(const uint64_t* __restrict__ offsets, uint32_t* __restrict__ keys,
                                   uint32_t* __restrict__ idx, size_t n_msgs) {
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < n_msgs; m += (size_t)gridDim.x * blockDim.x) {
        keys[m] = (uint32_t)((offsets[m + 1] - offsets[m]) / 4 + 1);
        idx[m] = (uint32_t)m;
    }
}

// ---- integer-multiply roofline microbenchmark --------------------------------------------------------
// 8 independent accumulators per thread, 8 unrolled steps per loop iteration; multiplicands come from
// the neighbouring accumulator so nothing is loop-invariant or warp-uniform (ptxas otherwise hoists the
// product and the loop degenerates into adds -- see tools/microbench.cu, which validates these forms and
// whose output is profiles/r01_microbench_pipe_costs.txt).  Products per thread per iteration: 8*8*P.
//   variant 0: 4-link carry chains (IMAD.WIDE.U32.X carry-in + carry-out), the production idiom; P = 4
//   variant 1: IMAD.WIDE.U32 with carry-out only + IADD3.X;                                      P = 1
//   variant 2: IMAD (32-bit low half only -- NOT a full product, context only);                  P = 1
//   variant 3: IMAD + IMAD.HI.U32 pair per product;                                              P = 1
//   variant 4: 16-link carry chains, 4 independent accumulators (products per iteration 8*4*16)
// 16-link carry chain (one landing add per 16 products): the closest a program gets to the bare pipe rate
__device__ __forceinline__ void cmad16(uint32_t (&acc)[33], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm(
        "mad.lo.cc.u32 %0, %33, %37, %0;\n\t"
        "madc.hi.cc.u32 %1, %33, %37, %1;\n\t"
        "madc.lo.cc.u32 %2, %34, %37, %2;\n\t"
        "madc.hi.cc.u32 %3, %34, %37, %3;\n\t"
        "madc.lo.cc.u32 %4, %35, %37, %4;\n\t"
        "madc.hi.cc.u32 %5, %35, %37, %5;\n\t"
        "madc.lo.cc.u32 %6, %36, %37, %6;\n\t"
        "madc.hi.cc.u32 %7, %36, %37, %7;\n\t"
        "madc.lo.cc.u32 %8, %33, %37, %8;\n\t"
        "madc.hi.cc.u32 %9, %33, %37, %9;\n\t"
        "madc.lo.cc.u32 %10, %34, %37, %10;\n\t"
        "madc.hi.cc.u32 %11, %34, %37, %11;\n\t"
        "madc.lo.cc.u32 %12, %35, %37, %12;\n\t"
        "madc.hi.cc.u32 %13, %35, %37, %13;\n\t"
        "madc.lo.cc.u32 %14, %36, %37, %14;\n\t"
        "madc.hi.cc.u32 %15, %36, %37, %15;\n\t"
        "madc.lo.cc.u32 %16, %33, %37, %16;\n\t"
        "madc.hi.cc.u32 %17, %33, %37, %17;\n\t"
        "madc.lo.cc.u32 %18, %34, %37, %18;\n\t"
        "madc.hi.cc.u32 %19, %34, %37, %19;\n\t"
        "madc.lo.cc.u32 %20, %35, %37, %20;\n\t"
        "madc.hi.cc.u32 %21, %35, %37, %21;\n\t"
        "madc.lo.cc.u32 %22, %36, %37, %22;\n\t"
        "madc.hi.cc.u32 %23, %36, %37, %23;\n\t"
        "madc.lo.cc.u32 %24, %33, %37, %24;\n\t"
        "madc.hi.cc.u32 %25, %33, %37, %25;\n\t"
        "madc.lo.cc.u32 %26, %34, %37, %26;\n\t"
        "madc.hi.cc.u32 %27, %34, %37, %27;\n\t"
        "madc.lo.cc.u32 %28, %35, %37, %28;\n\t"
        "madc.hi.cc.u32 %29, %35, %37, %29;\n\t"
        "madc.lo.cc.u32 %30, %36, %37, %30;\n\t"
        "madc.hi.cc.u32 %31, %36, %37, %31;\n\t"
        "addc.u32 %32, %32, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(acc[12]), "+r"(acc[13]), "+r"(acc[14]), "+r"(acc[15]), "+r"(acc[16]), "+r"(acc[17]), "+r"(acc[18]), "+r"(acc[19]), "+r"(acc[20]), "+r"(acc[21]), "+r"(acc[22]), "+r"(acc[23]), "+r"(acc[24]), "+r"(acc[25]), "+r"(acc[26]), "+r"(acc[27]), "+r"(acc[28]), "+r"(acc[29]), "+r"(acc[30]), "+r"(acc[31]), "+r"(acc[32])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

constexpr int kPeakIlp = 8;
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_peak_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                        int iters) {
    const uint32_t* my = in + (threadIdx.x & 63) * 32;
    uint32_t a = my[30], b = my[31], r = 0;
    if (VARIANT == 0) {
        uint32_t e[kPeakIlp][9];
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++)
#pragma unroll
            for (int q = 0; q < 9; q++) e[k][q] = my[(k + q) & 31];
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < kPeakIlp; k++) cmad4(e[k], a, b, a ^ 0x5555u, b ^ 0x3333u, e[(k + 1) & 7][1]);
        }
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++)
#pragma unroll
            for (int q = 0; q < 9; q++) r ^= e[k][q];
    } else if (VARIANT == 4) {
        uint32_t e[4][33];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int q = 0; q < 33; q++) e[k][q] = my[(k + q) & 31];
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < 4; k++) cmad16(e[k], a, b, a ^ 0x5555u, b ^ 0x3333u, e[(k + 1) & 3][1]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int q = 0; q < 33; q++) r ^= e[k][q];
    } else if (VARIANT == 1) {
        uint32_t lo[kPeakIlp], hi[kPeakIlp], t[kPeakIlp];
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) { lo[k] = my[k]; hi[k] = my[8 + k]; t[k] = 0; }
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < kPeakIlp; k++)
                    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                                 : "+r"(lo[k]), "+r"(hi[k]), "+r"(t[k]) : "r"(hi[(k + 1) & 7]), "r"(b));
        }
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) r ^= lo[k] ^ hi[k] ^ t[k];
    } else if (VARIANT == 2) {
        uint32_t acc[kPeakIlp];
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) acc[k] = my[k];
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < kPeakIlp; k++)
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 1) & 7]), "r"(b));
        }
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) r ^= acc[k];
    } else {
        uint32_t lo[kPeakIlp], hi[kPeakIlp];
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) { lo[k] = my[k]; hi[k] = my[8 + k]; }
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < kPeakIlp; k++)
                    asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;"
                                 : "+r"(lo[k]), "+r"(hi[k]) : "r"(hi[(k + 1) & 7]), "r"(b));
        }
#pragma unroll
        for (int k = 0; k < kPeakIlp; k++) r ^= lo[k] ^ hi[k];
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r ^ a;
}
constexpr double kPeakProductsPerIterV0 = 8.0 * 8.0 * 4.0;
constexpr double kPeakProductsPerIterV123 = 8.0 * 8.0;
constexpr double kPeakProductsPerIterV4 = 8.0 * 4.0 * 16.0;

}  // namespace hades
