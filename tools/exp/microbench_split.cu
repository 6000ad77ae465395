// One Montgomery product split over TWO lanes (lone-warp regime): lane L takes the low four limbs of b, lane H the
// high four; both run 4 multiply+reduce steps, then 4 reduce-only steps (only L's are used), one shuffle exchange and
// a 9-limb add: 80 links instead of 112 on the critical path.  Cycles per product against fr_mul_lazy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hades252_b200/csrc -o tools/exp/microbench_split tools/exp/microbench_split.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "fr.cuh"
using namespace hades;

template <int STEPS>
__device__ __forceinline__ void redc_short(uint32_t (&r)[9], const uint32_t (&t)[9]) {
    uint32_t A[9], B[9];
#pragma unroll
    for (int k = 0; k < 8; k++) { A[k] = t[k]; B[k] = 0; }
    A[8] = 0; B[8] = 0; B[7] = t[8];
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < STEPS; i += 2) {
        {
            uint32_t cx = (i == 0) ? 0u : add_carry_out(A[0], x);
            MontQ q = mont_quotient(A[0]);
            redc_odd(B, q); q.nz += cx; redc_even(A, q);
            x = A[1];
#pragma unroll
            for (int k = 0; k < 7; k++) A[k] = A[k + 2];
            A[7] = 0; A[8] = 0;
        }
        {
            uint32_t cx = add_carry_out(B[0], x);
            MontQ q = mont_quotient(B[0]);
            redc_odd(A, q); q.nz += cx; redc_even(B, q);
            x = B[1];
#pragma unroll
            for (int k = 0; k < 7; k++) B[k] = B[k + 2];
            B[7] = 0; B[8] = 0;
        }
    }
    merge_even_odd(r, A, B, x);
}

__device__ __forceinline__ void mmul_split(uint32_t (&r)[9], const uint32_t (&a)[8], const uint32_t (&b)[8], bool high) {
    uint32_t bb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) bb[k] = high ? b[4 + k] : b[k];
    uint32_t u[9], v[9];
    dot_mont_steps<1, 4>(u, [&](int, int k) { return a[k]; }, [&](int, int i) { return bb[i]; });
    redc_short<4>(v, u);
    uint32_t mine[9], other[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { mine[k] = high ? u[k] : v[k]; other[k] = __shfl_xor_sync(0xffffffffu, mine[k], 1); }
    asm("add.cc.u32 %0, %9, %18;\n\taddc.cc.u32 %1, %10, %19;\n\taddc.cc.u32 %2, %11, %20;\n\taddc.cc.u32 %3, %12, %21;\n\t"
        "addc.cc.u32 %4, %13, %22;\n\taddc.cc.u32 %5, %14, %23;\n\taddc.cc.u32 %6, %15, %24;\n\taddc.cc.u32 %7, %16, %25;\n\t"
        "addc.u32 %8, %17, %26;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8])
        : "r"(mine[0]), "r"(mine[1]), "r"(mine[2]), "r"(mine[3]), "r"(mine[4]), "r"(mine[5]), "r"(mine[6]), "r"(mine[7]), "r"(mine[8]),
          "r"(other[0]), "r"(other[1]), "r"(other[2]), "r"(other[3]), "r"(other[4]), "r"(other[5]), "r"(other[6]), "r"(other[7]), "r"(other[8]));
}

template <int IMPL>
__global__ void __launch_bounds__(32) k(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long* __restrict__ cyc, int iters) {
    const int pair = threadIdx.x >> 1;
    const bool high = threadIdx.x & 1;
    Fr x, y;
#pragma unroll
    for (int i = 0; i < 8; i++) { x.l[i] = in[pair * 16 + i]; y.l[i] = in[pair * 16 + 8 + i]; }
    x.l[7] &= 0x3fffffffu; y.l[7] &= 0x3fffffffu;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (IMPL == 0) {
            fr_mul(x, x, y);
        } else {
            uint32_t r[9];
            mmul_split(r, x.l, y.l, high);
            canon<1>(x, r);  // < 2p + ab/R < 4p
        }
    }
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; i++) out[(blockIdx.x * 32 + threadIdx.x) * 8 + i] = x.l[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static double median(std::vector<long long> v) { std::sort(v.begin(), v.end()); return (double)v[v.size() / 2]; }
int main() {
    const int sms = 148, iters = 2000;
    upload_modulus();
    std::vector<uint32_t> h(16 * 16);
    uint64_t s = 4242;
    for (auto& w : h) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; w = (uint32_t)(s >> 32); }
    uint32_t *din, *dout; long long* dcyc;
    cudaMalloc(&din, h.size() * 4); cudaMalloc(&dout, sms * 32 * 8 * 4); cudaMalloc(&dcyc, sms * 8);
    cudaMemcpy(din, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    std::vector<long long> c(sms);
    std::vector<uint32_t> r0(sms * 32 * 8), r1(sms * 32 * 8);
    k<0><<<sms, 32>>>(din, dout, dcyc, 10); k<0><<<sms, 32>>>(din, dout, dcyc, iters);
    cudaMemcpy(c.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost); cudaMemcpy(r0.data(), dout, r0.size() * 4, cudaMemcpyDeviceToHost);
    printf("fr_mul (one lane, canonical result)          %8.1f cycles per product\n", median(c) / iters);
    k<1><<<sms, 32>>>(din, dout, dcyc, 10); k<1><<<sms, 32>>>(din, dout, dcyc, iters);
    cudaMemcpy(c.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost); cudaMemcpy(r1.data(), dout, r1.size() * 4, cudaMemcpyDeviceToHost);
    printf("split over two lanes + exchange + canon<1>   %8.1f cycles per product\n", median(c) / iters);
    printf("results identical: %s   status: %s\n", r0 == r1 ? "yes" : "NO", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
