// What does ONE warp per scheduler cost?  (the small-batch kernels run in this regime.)
// One 32-thread block per SM, every pattern timed with clock64 inside the kernel (cycles per warp-instruction,
// median over the SMs).  Patterns: independent / carry-chained IMAD.WIDE, independent / dependent ALU, alternating
// and clustered mixes of the two, and the real Montgomery product of fr.cuh as one dependent chain and as two
// interleaved independent chains.  Check the SASS before trusting a line (cuobjdump -sass): ptxas reorders.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hades252_b200/csrc -o tools/microbench_lonewarp tools/microbench_lonewarp.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "fr.cuh"
using namespace hades;

#define UNROLL 16
template <int PATTERN>
__global__ void __launch_bounds__(32) k(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long* __restrict__ cyc, int iters) {
    uint32_t a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = in[threadIdx.x * 8 + i]; b[i] = in[256 + threadIdx.x * 8 + i] | 1; c[i] = in[512 + i]; }
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = ((uint64_t)a[i] << 32) | b[i];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (PATTERN == 0) {  // 8 independent IMAD.WIDE.U32 (64-bit accumulate, no carry predicates)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(b[i]), "r"(c[i]));
            } else if (PATTERN == 1) {  // one dependent IMAD.WIDE chain (latency)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(acc[0]) : "r"(c[i]));
            } else if (PATTERN == 2) {  // 8 independent IADD3-class adds
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c[i]));
            } else if (PATTERN == 3) {  // dependent add chain (ALU latency)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(c[i]));
            } else if (PATTERN == 4) {  // alternating: IMAD.WIDE, add (all independent)
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(b[i]), "r"(c[i]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c[i]));
                }
            } else if (PATTERN == 5) {  // alternating: IMAD.WIDE, 3 adds
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(b[i]), "r"(c[i]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c[i]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[(i + 3) & 7]) : "r"(c[i]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[(i + 5) & 7]) : "r"(c[i]));
                }
            } else if (PATTERN == 6) {  // clustered: 8 IMAD.WIDE then 8 adds
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(b[i]), "r"(c[i]));
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c[i]));
            } else if (PATTERN == 7) {  // the production carry-chain idiom: two independent accumulators, 4 links each
                uint32_t* e = a;
                asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\tmadc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                             "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\tmadc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                             : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7])
                             : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]));
                e = b;
                asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\tmadc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                             "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\tmadc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                             : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7])
                             : "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(c[1]));
            }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i] ^ (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
    out[blockIdx.x * 32 + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the real Montgomery product: NCHAIN independent dependent-chains of fr_mul per thread
template <int NCHAIN>
__global__ void __launch_bounds__(32) kmul(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long* __restrict__ cyc, int iters) {
    Fr x[NCHAIN], y;
#pragma unroll
    for (int s = 0; s < NCHAIN; s++)
#pragma unroll
        for (int i = 0; i < 8; i++) x[s].l[i] = in[threadIdx.x * 8 + i] ^ s;
#pragma unroll
    for (int i = 0; i < 8; i++) y.l[i] = in[256 + threadIdx.x * 8 + i];
#pragma unroll
    for (int s = 0; s < NCHAIN; s++) x[s].l[7] &= 0x3fffffffu;
    y.l[7] &= 0x3fffffffu;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int s = 0; s < NCHAIN; s++) fr_mul_lazy(x[s], x[s], y);
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int s = 0; s < NCHAIN; s++)
#pragma unroll
        for (int i = 0; i < 8; i++) r ^= x[s].l[i];
    out[blockIdx.x * 32 + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double median(std::vector<long long> v) { std::sort(v.begin(), v.end()); return (double)v[v.size() / 2]; }

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    upload_modulus();
    std::vector<uint32_t> h(1024);
    uint64_t s = 99;
    for (auto& w : h) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; w = (uint32_t)(s >> 32); }
    uint32_t *din, *dout;
    long long* dcyc;
    cudaMalloc(&din, 4096); cudaMalloc(&dout, sms * 32 * 4); cudaMalloc(&dcyc, sms * 8);
    cudaMemcpy(din, h.data(), 4096, cudaMemcpyHostToDevice);
    std::vector<long long> c(sms);
    printf("%s: ONE warp per SM (so one warp on its scheduler), cycles from clock64, median over %d SMs\n", prop.name, sms);
    const int iters = 200;
#define RUN(P, NAME, PER_TRIP, WHAT)                                                                     \
    k<P><<<sms, 32>>>(din, dout, dcyc, 10); k<P><<<sms, 32>>>(din, dout, dcyc, iters);                   \
    cudaMemcpy(c.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost);                                         \
    printf("%-52s %8.2f cycles per %s\n", NAME, median(c) / ((double)iters * UNROLL * PER_TRIP), WHAT);
    RUN(0, "IMAD.WIDE, 8 independent accumulators", 8, "IMAD.WIDE")
    RUN(1, "IMAD.WIDE, one dependent chain", 8, "IMAD.WIDE (latency)")
    RUN(2, "IADD, 8 independent", 8, "IADD")
    RUN(3, "IADD, one dependent chain", 8, "IADD (latency)")
    RUN(4, "alternating IMAD.WIDE + 1 IADD (independent)", 8, "pair")
    RUN(5, "alternating IMAD.WIDE + 3 IADD (independent)", 8, "group of 4")
    RUN(6, "clustered 8 IMAD.WIDE then 8 IADD", 8, "pair")
    RUN(7, "carry chains: 2 accumulators x 4 IMAD.WIDE.X links", 8, "IMAD.WIDE.X")
#define RUNMUL(N, NAME)                                                                                  \
    kmul<N><<<sms, 32>>>(din, dout, dcyc, 10); kmul<N><<<sms, 32>>>(din, dout, dcyc, 2000);              \
    cudaMemcpy(c.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost);                                         \
    printf("%-52s %8.1f cycles per Montgomery product (112 IMAD.WIDE + ~100 ALU)\n", NAME, median(c) / (2000.0 * N));
    RUNMUL(1, "fr_mul_lazy, one dependent chain")
    RUNMUL(2, "fr_mul_lazy, 2 independent chains interleaved")
    RUNMUL(4, "fr_mul_lazy, 4 independent chains interleaved")
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
