// Product-scanning Montgomery product (IMAD.WIDE .. RZ + 3-input adds on the ALU pipe) against the
// production carry-chain product of fr.cuh: cycles per product per SM sub-partition, saturated
// (128 x 5 blocks per SM, the production launch shape) and with one warp per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hades252_b200/csrc -o tools/exp/microbench_ps tools/exp/microbench_ps.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "fr.cuh"
#include "ps.cuh"
using namespace hades;

template <int IMPL, int TPB, int BPS>
__global__ void __launch_bounds__(TPB, BPS) k(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = in[(tid & 1023) * 16 + i]; b[i] = in[(tid & 1023) * 16 + 8 + i]; }
    a[7] &= 0x3fffffffu; b[7] &= 0x3fffffffu;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (IMPL == 0) {
            Fr x, y, z;
#pragma unroll
            for (int i = 0; i < 8; i++) { x.l[i] = a[i]; y.l[i] = b[i]; }
            fr_mul_lazy(z, x, y);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = z.l[i];
        } else {
            uint32_t r[9];
            ps::mul_mont(r, a, b, c_modp, c_modp[8]);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = r[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[(size_t)tid * 8 + i] = a[i];
}

template <int IMPL, int TPB, int BPS>
static double run(const char* name, int blocks, const uint32_t* din, uint32_t* dout, int iters, std::vector<uint32_t>& res) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<IMPL, TPB, BPS><<<blocks, TPB>>>(din, dout, 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<IMPL, TPB, BPS><<<blocks, TPB>>>(din, dout, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    res.resize((size_t)blocks * TPB * 8);
    cudaMemcpy(res.data(), dout, res.size() * 4, cudaMemcpyDeviceToHost);
    double warps_per_smsp = (double)blocks * TPB / 32.0 / (148.0 * 4.0);
    if (warps_per_smsp < 1) warps_per_smsp = 1;
    double cyc = best * 1e-3 * 1.965e9 / ((double)iters * warps_per_smsp);
    printf("%-44s %9.3f ms  %8.1f cycles per product per scheduler (%.1f warps/scheduler)  [%s]\n", name, best, cyc, warps_per_smsp,
           cudaGetErrorString(cudaGetLastError()));
    return cyc;
}

int main() {
    upload_modulus();
    std::vector<uint32_t> h(1024 * 16);
    uint64_t s = 12345;
    for (auto& w : h) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; w = (uint32_t)(s >> 32); }
    uint32_t *din, *dout;
    cudaMalloc(&din, h.size() * 4); cudaMalloc(&dout, (size_t)148 * 5 * 128 * 8 * 4);
    cudaMemcpy(din, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    std::vector<uint32_t> r0, r1;
    const int iters = 4000;
    run<0, 128, 5>("carry chains (fr.cuh), 128x5 per SM", 148 * 5, din, dout, iters, r0);
    run<1, 128, 5>("product scanning, 128x5 per SM", 148 * 5, din, dout, iters, r1);
    printf("results identical: %s\n", r0 == r1 ? "yes" : "NO");
    run<0, 128, 5>("carry chains, 128x2 per SM", 148 * 2, din, dout, iters, r0);
    run<1, 128, 5>("product scanning, 128x2 per SM", 148 * 2, din, dout, iters, r1);
    run<0, 128, 5>("carry chains, one warp per scheduler", 148, din, dout, iters, r0);
    run<1, 128, 5>("product scanning, one warp per scheduler", 148, din, dout, iters, r1);
    printf("results identical: %s\n", r0 == r1 ? "yes" : "NO");
    run<0, 256, 4>("carry chains, 256x4 per SM (64 regs)", 148 * 4, din, dout, iters, r0);
    run<1, 256, 4>("product scanning, 256x4 per SM (64 regs)", 148 * 4, din, dout, iters, r1);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
