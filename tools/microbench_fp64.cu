// Does the FP64 pipe pay for the Hades252 field arithmetic on B200?  (measure, do not argue.)
//
// DFMA issues every 2.2 cycles per scheduler and co-issues with a saturated IMAD.WIDE chain
// (profiles/r01_microbench_pipe_costs.txt), so a Montgomery product in 5 x 52-bit limbs carried in doubles is a
// candidate second multiplier next to the 8 x 32-bit IMAD.WIDE product the kernels use (fr.cuh, 112 products).
// This file measures, on the same launch shape as the production kernel (128-thread blocks, 5 per SM):
//   imad      fr_mul of fr.cuh (dot_mont<1> + conditional subtraction), dependent chain, ILP streams per thread
//   dfma      the same chain with a Montgomery product built from DFMA hi/lo product splitting:
//               hi = fma_rz(a, b, 2^104); lo = fma_rz(a, b, (2^104 + 2^52) - hi)   (3 FP64 ops per 52x52 product)
//             accumulated as raw bit patterns in 64-bit integer columns (2 integer adds per product), quotient digits
//             from the special form of p mod 2^52 (shifts and adds), R = 2^260, result < 1.2 p without a final
//             subtraction
//   hybrid    NI imad products and ND dfma products per loop trip in the SAME thread (independent streams): does the
//             second pipe come for free, as the DFMA-next-to-IMAD microbenchmark suggested?
// Output: cycles per Montgomery product per scheduler (SM sub-partition) for each variant, at the SM clock read from
// the device.  A few products are printed for an offline big-integer check (tools/check_fp64_mul.py).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hades252_b200/csrc -o microbench_fp64 tools/microbench_fp64.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fr.cuh"

using namespace hades;

// ---- 5 x 52-bit Montgomery product on the FP64 pipe ------------------------------------------------------------
struct F52 {
    double l[5];  // exact integers in [0, 2^52) (the top limb may carry a little more: values stay below 2^260)
};

__constant__ double c_p52[5];  // the modulus in 52-bit limbs, as doubles

__device__ __forceinline__ long long dbits(double x) { return __double_as_longlong(x); }

// c[k] += lo part of a*b, c[k+1] += hi part, as raw bit patterns (the exponent biases are removed once per column)
__device__ __forceinline__ void mad52(long long& clo, long long& chi, double a, double b) {
    const double c1 = 20282409603651670423947251286016.0;   // 2^104
    const double c2 = 20282409603651674927546878656512.0;   // 2^104 + 2^52
    const double hi = __fma_rz(a, b, c1);
    const double sub = __dadd_rn(c2, -hi);
    const double lo = __fma_rz(a, b, sub);
    clo += dbits(lo);
    chi += dbits(hi);
}

__device__ __forceinline__ void mul52(F52& r, const F52& a, const F52& b) {
    const long long kB52 = 0x4330000000000000LL;    // bits of 2^52
    const long long kB104 = 0x4670000000000000LL;   // bits of 2^104
    const long long kMask = (1LL << 52) - 1;
    long long c[11];
    // column k receives (number of lo parts) biases of 2^52 and (number of hi parts) biases of 2^104
#pragma unroll
    for (int k = 0; k < 11; k++) {
        int nlo = 0, nhi = 0;
#pragma unroll
        for (int i = 0; i < 5; i++)
#pragma unroll
            for (int j = 0; j < 5; j++) {
                if (i + j == k) nlo++;
                if (i + j + 1 == k) nhi++;
            }
        // products a*b and q*p land on the same columns shifted by the reduction step: q_i * p_j -> column i + j
        c[k] = -(long long)(2 * nlo) * kB52 - (long long)(2 * nhi) * kB104;
    }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) mad52(c[i + j], c[i + j + 1], a.l[i], b.l[j]);
    // Montgomery reduction, one 52-bit digit per step; -1/p mod 2^52 = -(1 + 2^32) because p = 1 - 2^32 (mod 2^52)
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const long long t = c[i] & kMask;
        const long long q = (-(t + (t << 32))) & kMask;
        const double qd = __dadd_rn(__longlong_as_double(kB52 | q), -4503599627370496.0);
#pragma unroll
        for (int j = 0; j < 5; j++) mad52(c[i + j], c[i + j + 1], qd, c_p52[j]);
        c[i + 1] += c[i] >> 52;  // the low 52 bits are zero now
    }
    // carry-normalise the upper half and go back to doubles
#pragma unroll
    for (int k = 5; k < 9; k++) {
        c[k + 1] += c[k] >> 52;
        c[k] &= kMask;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) r.l[k] = __dadd_rn(__longlong_as_double(kB52 | c[5 + k]), -4503599627370496.0);
}

// ---- kernels ---------------------------------------------------------------------------------------------------
constexpr int kThreads = 128;

template <int NI, int ND>
__global__ void __launch_bounds__(kThreads, 5) chain_kernel(const uint32_t* __restrict__ in32, const double* __restrict__ in52,
                                                            uint32_t* __restrict__ out32, double* __restrict__ out52, int iters) {
    const int t = blockIdx.x * kThreads + threadIdx.x;
    Fr x[NI > 0 ? NI : 1], y[NI > 0 ? NI : 1];
    F52 u[ND > 0 ? ND : 1], v[ND > 0 ? ND : 1];
#pragma unroll
    for (int s = 0; s < NI; s++)
#pragma unroll
        for (int k = 0; k < 8; k++) { x[s].l[k] = in32[((t * 2 + 0) % 4096) * 8 + k] ^ (uint32_t)s; y[s].l[k] = in32[((t * 2 + 1) % 4096) * 8 + k]; }
#pragma unroll
    for (int s = 0; s < NI; s++) { x[s].l[7] &= 0x3fffffffu; y[s].l[7] &= 0x3fffffffu; }
#pragma unroll
    for (int s = 0; s < ND; s++)
#pragma unroll
        for (int k = 0; k < 5; k++) { u[s].l[k] = in52[((t * 2 + 0) % 4096) * 5 + k] + (k == 0 ? (double)s : 0.0); v[s].l[k] = in52[((t * 2 + 1) % 4096) * 5 + k]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < NI; s++) fr_mul(x[s], x[s], y[s]);
#pragma unroll
        for (int s = 0; s < ND; s++) mul52(u[s], u[s], v[s]);
    }
    if (NI > 0) {
        uint32_t acc = 0;
#pragma unroll
        for (int s = 0; s < NI; s++)
#pragma unroll
            for (int k = 0; k < 8; k++) acc ^= x[s].l[k];
        out32[t] = acc;
    }
    if (ND > 0) {
        double acc = 0;
#pragma unroll
        for (int s = 0; s < ND; s++)
#pragma unroll
            for (int k = 0; k < 5; k++) acc += u[s].l[k];
        out52[t] = acc;
    }
}

// one product per thread, operands and result stored: the offline check
__global__ void check_kernel(const double* __restrict__ in52, double* __restrict__ out, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    F52 a, b, r;
    for (int k = 0; k < 5; k++) { a.l[k] = in52[(2 * t) * 5 + k]; b.l[k] = in52[(2 * t + 1) * 5 + k]; }
    mul52(r, a, b);
    mul52(r, r, b);  // a second product on a non-normalised input (< 1.2 p, top limb not masked)
    for (int k = 0; k < 5; k++) out[t * 5 + k] = r.l[k];
}

static uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

template <int NI, int ND>
static double run(const char* name, const uint32_t* d32, const double* d52, uint32_t* o32, double* o52, double mhz, int sms, double* base_i, double* base_d) {
    const int blocks = sms * 5 * 4, iters = 2000;
    chain_kernel<NI, ND><<<blocks, kThreads>>>(d32, d52, o32, o52, 50);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        chain_kernel<NI, ND><<<blocks, kThreads>>>(d32, d52, o32, o52, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaFuncAttributes attr;
    cudaFuncGetAttributes(&attr, chain_kernel<NI, ND>);
    // warp-level trips per scheduler
    const double warps_per_sched = (double)blocks * (kThreads / 32) / (sms * 4);
    const double cycles_per_trip = best * 1e-3 * mhz * 1e6 / (warps_per_sched * iters);
    double equiv = 0;  // time of the trip expressed in IMAD-only products: what the hybrid has to beat
    if (NI > 0 && ND == 0) *base_i = cycles_per_trip / NI;
    if (ND > 0 && NI == 0) *base_d = cycles_per_trip / ND;
    printf("%-34s %3d regs %4zu B stack  %8.3f ms  %9.1f cycles/trip/scheduler  %8.1f cycles per product", name, attr.numRegs,
           (size_t)attr.localSizeBytes, best, cycles_per_trip, cycles_per_trip / (NI + ND));
    if (NI > 0 && ND > 0 && *base_i > 0) {
        equiv = (NI + ND) * *base_i;
        printf("   (IMAD-only for the same %d products: %.1f -> hybrid speed-up %.3fx)", NI + ND, equiv, equiv / cycles_per_trip);
    }
    printf("\n");
    return cycles_per_trip;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, SM clock %.0f MHz (max); launch: 128-thread blocks, 5 per SM (the production shape)\n", prop.name,
           prop.multiProcessorCount, mhz);
    upload_modulus();
    const uint64_t P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    double p52[5];
    {
        unsigned __int128 lo = ((unsigned __int128)P[1] << 64) | P[0], hi = ((unsigned __int128)P[3] << 64) | P[2];
        auto bits = [&](int from) -> uint64_t {  // 52 bits of the 256-bit p starting at bit `from`
            uint64_t r = 0;
            for (int b = 0; b < 52; b++) {
                int pos = from + b;
                uint64_t bit = pos < 128 ? (uint64_t)(lo >> pos) & 1 : pos < 256 ? (uint64_t)(hi >> (pos - 128)) & 1 : 0;
                r |= bit << b;
            }
            return r;
        };
        for (int k = 0; k < 5; k++) p52[k] = (double)bits(52 * k);
    }
    cudaMemcpyToSymbol(c_p52, p52, sizeof p52);
    uint64_t seed = 42;
    std::vector<uint32_t> h32(4096 * 8);
    std::vector<double> h52(4096 * 5);
    for (auto& w : h32) w = (uint32_t)splitmix(seed);
    for (size_t i = 0; i < h52.size(); i++) h52[i] = (double)(splitmix(seed) & ((i % 5 == 4) ? ((1ULL << 46) - 1) : ((1ULL << 52) - 1)));  // < 2^254
    uint32_t *d32, *o32;
    double *d52, *o52;
    const int threads_total = prop.multiProcessorCount * 5 * 4 * kThreads;
    cudaMalloc(&d32, h32.size() * 4); cudaMalloc(&d52, h52.size() * 8);
    cudaMalloc(&o32, threads_total * 4); cudaMalloc(&o52, threads_total * 8);
    cudaMemcpy(d32, h32.data(), h32.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d52, h52.data(), h52.size() * 8, cudaMemcpyHostToDevice);
    // offline check material: 64 products
    {
        double* dchk;
        cudaMalloc(&dchk, 64 * 5 * 8);
        check_kernel<<<1, 64>>>(d52, dchk, 64);
        std::vector<double> r(64 * 5);
        cudaMemcpy(r.data(), dchk, r.size() * 8, cudaMemcpyDeviceToHost);
        for (int t = 0; t < 64; t++) {
            printf("CHECK");
            for (int k = 0; k < 5; k++) printf(" %.0f", h52[(2 * t) * 5 + k]);
            for (int k = 0; k < 5; k++) printf(" %.0f", h52[(2 * t + 1) * 5 + k]);
            for (int k = 0; k < 5; k++) printf(" %.0f", r[t * 5 + k]);
            printf("\n");
        }
    }
    const int sms = prop.multiProcessorCount;
    double bi = 0, bd = 0;
    run<1, 0>("imad  x1 (fr_mul, 8x32 IMAD.WIDE)", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    run<2, 0>("imad  x2", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    run<0, 1>("dfma  x1 (5x52, DFMA hi/lo split)", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    run<0, 2>("dfma  x2", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    run<2, 1>("hybrid 2 imad + 1 dfma", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    run<3, 1>("hybrid 3 imad + 1 dfma", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    run<4, 1>("hybrid 4 imad + 1 dfma", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    run<1, 1>("hybrid 1 imad + 1 dfma", d32, d52, o32, o52, mhz, sms, &bi, &bd);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
