// Test-only kernels: the device field arithmetic of fr.cuh on caller-supplied operands, so that the real PTX
// carry chains (mad.lo.cc / madc.hi.cc, the special-modulus reduction steps, the 512-bit addend injection) are
// checked directly against big-integer arithmetic and not only through whole permutations.  Reference call sites
// of the operations: src/strategies/scalar.rs:28 (`+=`), :33 (`square`, `*`), :44 (`*`, `+=`).
// Not part of the reference-facing surface (include/hades_cuda.h: measurement / test helpers).
#include <cuda_runtime.h>

#include <mutex>
#include <set>

#include "fr.cuh"
#include "width_ops.hpp"

namespace hades {
namespace {

template <int OP>
struct OpShape;  // words in / out per element
#define HADES_OP(op, in, out) template <> struct OpShape<op> { static constexpr int kIn = in, kOut = out; }
HADES_OP(0, 16, 8);   // fr_mul(a, b)
HADES_OP(1, 16, 8);   // fr_add(a, b)
HADES_OP(2, 8, 8);    // fr_sbox(x)
HADES_OP(3, 8, 9);    // sqr_mont(a), raw 9 limbs
HADES_OP(4, 16, 16);  // mul_wide(a, b)
HADES_OP(5, 16, 9);   // redc16(t), raw
HADES_OP(6, 64, 9);   // dot_mont<4>(A[4], B[4]), raw
HADES_OP(7, 80, 9);   // dot_mont_plus<4>(A[4], B[4], t[16]), raw
HADES_OP(8, 80, 9);   // dot_mont<5>(A[5], B[5]), raw
HADES_OP(9, 16, 9);   // dot_mont<1>(a, b), raw (the cooperative kernel's slot product)
HADES_OP(10, 9, 8);   // canon<0>
HADES_OP(11, 9, 8);   // canon<1>
HADES_OP(12, 9, 8);   // canon<2>
HADES_OP(13, 9, 8);   // canon<3>
HADES_OP(14, 9, 8);   // canon<4>
HADES_OP(15, 64, 9);  // mul_const_short<4>(X[4], y): in = X_0..X_3 (32 words), y (8), padding
HADES_OP(16, 16, 8);  // fr_mul_lazy(a, b): 8 limbs, not canonical
constexpr int kNumOps = 17;

template <int OP>
__global__ void __launch_bounds__(128) fr_op_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    constexpr int kIn = OpShape<OP>::kIn, kOut = OpShape<OP>::kOut;
    uint32_t x[kIn], y[kOut];
#pragma unroll
    for (int k = 0; k < kIn; k++) x[k] = in[i * kIn + k];
    if constexpr (OP == 0 || OP == 1 || OP == 16) {
        Fr a, b, r;
#pragma unroll
        for (int k = 0; k < 8; k++) { a.l[k] = x[k]; b.l[k] = x[8 + k]; }
        if constexpr (OP == 0) fr_mul(r, a, b);
        else if constexpr (OP == 1) fr_add(r, a, b);
        else fr_mul_lazy(r, a, b);
#pragma unroll
        for (int k = 0; k < 8; k++) y[k] = r.l[k];
    } else if constexpr (OP == 2) {
        Fr a;
#pragma unroll
        for (int k = 0; k < 8; k++) a.l[k] = x[k];
        fr_sbox(a);
#pragma unroll
        for (int k = 0; k < 8; k++) y[k] = a.l[k];
    } else if constexpr (OP == 3) {
        uint32_t a[8], r[9];
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = x[k];
        sqr_mont(r, a);
#pragma unroll
        for (int k = 0; k < 9; k++) y[k] = r[k];
    } else if constexpr (OP == 4) {
        uint32_t a[8], b[8], t[16];
#pragma unroll
        for (int k = 0; k < 8; k++) { a[k] = x[k]; b[k] = x[8 + k]; }
        mul_wide(t, a, b);
#pragma unroll
        for (int k = 0; k < 16; k++) y[k] = t[k];
    } else if constexpr (OP == 5) {
        uint32_t t[16], r[9];
#pragma unroll
        for (int k = 0; k < 16; k++) t[k] = x[k];
        redc16(r, t);
#pragma unroll
        for (int k = 0; k < 9; k++) y[k] = r[k];
    } else if constexpr (OP == 6 || OP == 8 || OP == 9) {
        constexpr int N = OP == 6 ? 4 : OP == 8 ? 5 : 1;
        uint32_t r[9];
        dot_mont<N>(r, [&](int j, int k) { return x[8 * j + k]; }, [&](int j, int i2) { return x[8 * N + 8 * j + i2]; });
#pragma unroll
        for (int k = 0; k < 9; k++) y[k] = r[k];
    } else if constexpr (OP == 7) {
        uint32_t r[9], t[16];
#pragma unroll
        for (int k = 0; k < 16; k++) t[k] = x[64 + k];
        dot_mont_plus<4>(r, [&](int j, int k) { return x[8 * j + k]; }, [&](int j, int i2) { return x[32 + 8 * j + i2]; }, t);
#pragma unroll
        for (int k = 0; k < 9; k++) y[k] = r[k];
    } else if constexpr (OP >= 10 && OP <= 14) {
        uint32_t r[9];
#pragma unroll
        for (int k = 0; k < 9; k++) r[k] = x[k];
        Fr o;
        canon<OP - 10>(o, r);
#pragma unroll
        for (int k = 0; k < 8; k++) y[k] = o.l[k];
    } else if constexpr (OP == 15) {
        uint32_t r[9];
        Fr v;
#pragma unroll
        for (int k = 0; k < 8; k++) v.l[k] = x[32 + k];
        mul_const_short<4>(r, [&](int j, int k) { return x[8 * j + k]; }, v);
#pragma unroll
        for (int k = 0; k < 9; k++) y[k] = r[k];
    }
#pragma unroll
    for (int k = 0; k < kOut; k++) out[i * kOut + k] = y[k];
}

std::mutex g_mutex;
std::set<int> g_uploaded;  // devices whose copy of this translation unit's modulus table is resident

template <int OP>
cudaError_t launch(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t s) {
    fr_op_kernel<OP><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(in, out, n);
    return cudaGetLastError();
}

}  // namespace

// shapes: 0 on success
int fr_test_shape(int op, int* in_words, int* out_words) {
#define HADES_SHAPE(o) case o: *in_words = OpShape<o>::kIn; *out_words = OpShape<o>::kOut; return 0;
    switch (op) {
        HADES_SHAPE(0) HADES_SHAPE(1) HADES_SHAPE(2) HADES_SHAPE(3) HADES_SHAPE(4) HADES_SHAPE(5) HADES_SHAPE(6) HADES_SHAPE(7)
        HADES_SHAPE(8) HADES_SHAPE(9) HADES_SHAPE(10) HADES_SHAPE(11) HADES_SHAPE(12) HADES_SHAPE(13) HADES_SHAPE(14)
        HADES_SHAPE(15) HADES_SHAPE(16)
        default: return 1;
    }
}

cudaError_t fr_test_launch(int op, const uint32_t* d_in, uint32_t* d_out, size_t n, cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        if (!g_uploaded.count(dev)) {
            e = upload_modulus();
            if (e != cudaSuccess) return e;
            g_uploaded.insert(dev);
        }
    }
    if (n == 0) return cudaSuccess;
#define HADES_LAUNCH(o) case o: return launch<o>(d_in, d_out, n, s);
    switch (op) {
        HADES_LAUNCH(0) HADES_LAUNCH(1) HADES_LAUNCH(2) HADES_LAUNCH(3) HADES_LAUNCH(4) HADES_LAUNCH(5) HADES_LAUNCH(6) HADES_LAUNCH(7)
        HADES_LAUNCH(8) HADES_LAUNCH(9) HADES_LAUNCH(10) HADES_LAUNCH(11) HADES_LAUNCH(12) HADES_LAUNCH(13) HADES_LAUNCH(14)
        HADES_LAUNCH(15) HADES_LAUNCH(16)
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hades
