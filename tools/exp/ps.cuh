// product-scanning Montgomery product, radix 2^32, products without addend (IMAD.WIDE .. RZ) and
// column sums on the ALU pipe through 3-input adds
#pragma once
#include <stdint.h>
namespace ps {
__device__ __forceinline__ uint32_t lo32(uint64_t x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }

struct Col {
    uint64_t L, H;   // L: sum of the 32-bit values of this column; H: of the next one
    uint32_t pl, ph; bool have;
    __device__ __forceinline__ Col() : L(0), H(0), pl(0), ph(0), have(false) {}
    __device__ __forceinline__ void prod(uint32_t a, uint32_t b) {
        uint64_t r = (uint64_t)a * b;
        if (have) { L = L + (uint64_t)pl + (uint64_t)lo32(r); H = H + (uint64_t)ph + (uint64_t)hi32(r); have = false; }
        else { pl = lo32(r); ph = hi32(r); have = true; }
    }
    __device__ __forceinline__ void val(uint32_t v) {  // a 32-bit value of this column
        if (have) { L = L + (uint64_t)pl + (uint64_t)v; H = H + (uint64_t)ph; have = false; }
        else L = L + (uint64_t)v;
    }
    __device__ __forceinline__ void flush() { if (have) { L += pl; H += ph; have = false; } }
    __device__ __forceinline__ uint32_t low() { flush(); return lo32(L); }
    __device__ __forceinline__ void next() { L = H + (L >> 32); H = 0; }
};

// r = a*b/2^256 mod p (9th limb returned), p limbs from pm[] (opaque)
__device__ __forceinline__ void mul_mont(uint32_t (&r)[9], const uint32_t (&a)[8], const uint32_t (&b)[8], const uint32_t* pm, uint32_t zero) {
    Col c;
    uint32_t m[8], t[8], nz[8];
#pragma unroll
    for (int k = 0; k < 16; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { int j = k - i; if (j >= 0 && j < 8) c.prod(a[i], b[j]); }
#pragma unroll
        for (int i = 0; i < 8; i++) { int j = k - i; if (i < k && j >= 2 && j < 8) c.prod(m[i], pm[j]); }
        if (k >= 1 && k - 1 < 8) { c.val(t[k - 1]); c.val(nz[k - 1]); }
        if (k >= 2 && k - 2 < 8) c.val(m[k - 2] - nz[k - 2]);
        uint32_t w = c.low();
        if (k < 8) { t[k] = w; m[k] = zero - w; nz[k] = w ? 1u : 0u; }
        else r[k - 8] = w;
        c.next();
    }
    r[8] = lo32(c.L);
}
}
