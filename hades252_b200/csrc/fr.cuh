// BLS12-381 scalar field (Fr) arithmetic for sm_100a, 8 x 32-bit limbs, Montgomery form R = 2^256.
//
// Device replacement for the `BlsScalar` operations on the reference's hot path (reference call
// sites: src/strategies/scalar.rs:28 `+=`, :33 `square`/`*`, :44 `*`,`+=`; the arithmetic itself
// is the external crate dusk-bls12_381 0.13, Cargo.toml:12).  The in-memory format equals
// `BlsScalar` (4 LE u64 Montgomery limbs == 8 LE u32 limbs), so states cross the FFI without
// conversion, and every public result is fully reduced to [0,p): representations are unique, so
// results are bit-identical to the reference's.
//
// Design (B200 / sm_100a):
//  * every 32x32->64 product is a PTX `mad.lo.cc.u32` + `madc.hi.cc.u32` pair inside one carry
//    chain; ptxas fuses each pair into ONE `IMAD.WIDE.U32.X Rd, Pout, Ra, Rb|UR, Rc, Pin` (carry
//    in a predicate), so the int-multiply pipe sees exactly one op per limb product;
//  * "even/odd" accumulators: products a[k]*b with k even land on 64-bit columns of `even`
//    (limb positions 0..8), k odd on columns of `odd` (positions 1..9): no carry ripples;
//  * word-serial Montgomery reduction interleaved with the products (CIOS), generalised to an
//    N-term dot product  sum_j c_j*v_j  that is reduced ONCE (lazy reduction of an MDS row);
//  * special modulus: p[0] = 1 and -p^{-1} mod 2^32 = 0xffffffff, so the Montgomery quotient is
//    m = -t0 (no multiply); m*p[0] and m*p[1] (p[1] = 2^32 - 1) need no product: 6 products and 8
//    ALU instructions per reduction step (mont_quotient, redc_odd, redc_even);
//  * a 512-bit addend can share a dot product's reduction at no cost (dot_mont_plus), which lets the
//    partial rounds keep the S-box output x^4 * x as an unreduced product (mul_wide).
//  Cost model measured on B200: 4.0 cycles per IMAD.WIDE and ~0.63 per other
//  instruction per SM sub-partition -- ALU instructions are not free beside the multiply pipe.
//
// The same source compiles with g++ -DHADES_HOST_EMUL (tests/host_emul): the PTX chains are then
// replaced by portable C++ of identical semantics, so the limb-level algorithm is checked against
// the CPU oracle without a GPU.
#pragma once
#include <stdint.h>

#if defined(HADES_HOST_EMUL)
#include <assert.h>
#define HADES_DEV inline
#define HADES_EMUL 1
#define HADES_ASSERT(x) assert(x)
#else
#define HADES_DEV __device__ __forceinline__
#define HADES_EMUL 0
#define HADES_ASSERT(x)
#endif

namespace hades {

// p = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001 (README.md:35)
// limb k (0..8) of p << s; every use has compile-time k, s, so these fold to immediates.
HADES_DEV constexpr uint32_t p_limb(int k) {
    return k == 0 ? 0x00000001u : k == 1 ? 0xffffffffu : k == 2 ? 0xfffe5bfeu : k == 3 ? 0x53bda402u
         : k == 4 ? 0x09a1d805u : k == 5 ? 0x3339d808u : k == 6 ? 0x299d7d48u : k == 7 ? 0x73eda753u : 0u;
}
HADES_DEV constexpr uint32_t p_shl_limb(int s, int k) {
    return s == 0 ? p_limb(k) : ((p_limb(k) << s) | (k > 0 ? (p_limb(k - 1) >> (32 - s)) : 0u));
}

struct Fr {
    uint32_t l[8];
};

// Modulus limbs as MULTIPLICANDS.  On the device they are read from `__constant__` memory so that
// ptxas sees an opaque uniform operand: with immediates it special-cases p[1] = 0xffffffff into
// IMAD.HI.U32 + IADD3 and splits many IMAD.WIDE.U32.X into IMAD.X + IMAD.HI.U32.X pairs, which cost
// 2.0 + 5.1 pipe cycles instead of 4.05 (profiles/r01_microbench_pipe_costs.txt).
// Deliberately NOT statically initialised (an initialiser gets folded back into immediates): every
// translation unit that multiplies by the modulus uploads it once with upload_modulus().
#if !HADES_EMUL
static __constant__ uint32_t c_modp[9];  // p[0..7], then a zero the compiler cannot see (mont_quotient)
HADES_DEV uint32_t modp(int k) { return c_modp[k]; }
#if defined(__CUDACC__)
static inline cudaError_t upload_modulus() {
    const uint32_t p[9] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u, 0u};
    return cudaMemcpyToSymbol(c_modp, p, sizeof(p), 0, cudaMemcpyHostToDevice);
}
#endif
#else
HADES_DEV uint32_t modp(int k) { return p_limb(k); }
#endif

// Montgomery quotient of one reduction step: m = -t0 mod 2^32 (because -p^-1 = -1 mod 2^32), and
// nz = (t0 != 0) = the carry of t0 + m.  m is computed as  c - t0  with c a zero read from constant memory:
// if ptxas can prove m == -t0 it rewrites every m*p[k] product into IMAD.X + IMAD.HI.U32.X pairs (7.1 pipe
// cycles instead of 4.05).  Three ALU instructions per step: m, nz, m - nz.
struct MontQ {
    uint32_t t0, m, nz, mm;
};
HADES_DEV MontQ mont_quotient(uint32_t t0) {
    MontQ q;
    q.t0 = t0;
    q.m = modp(8) - t0;
#if !HADES_EMUL
    asm("min.u32 %0, %1, 1;" : "=r"(q.nz) : "r"(t0));  // one VIMNMX (a ternary compiles to ISETP + SEL)
#else
    q.nz = t0 ? 1u : 0u;
#endif
    q.mm = q.m - q.nz;
    return q;
}

// ------------------------------------------------------------------------------------------------
// Chain primitives.  Each is ONE asm statement: the carry flag never lives across statements.
// ------------------------------------------------------------------------------------------------
#if HADES_EMUL
namespace emul {
// acc (64-bit column k = limbs 2k,2k+1) += a[k]*b, k = k0..3, with carry-in; returns carry-out.
inline uint64_t chain(uint32_t* acc, const uint32_t* a, int k0, uint32_t b, uint64_t carry) {
    for (int k = k0; k < 4; k++) {
        unsigned __int128 t = (unsigned __int128)a[k] * b + (((uint64_t)acc[2 * k + 1] << 32) | acc[2 * k]) + carry;
        acc[2 * k] = (uint32_t)t;
        acc[2 * k + 1] = (uint32_t)(t >> 32);
        carry = (uint64_t)(t >> 64);
    }
    return carry;
}
inline void top(uint32_t& t, uint64_t carry) {
    uint64_t s = (uint64_t)t + carry;
    HADES_ASSERT((s >> 32) == 0);  // the 9th limb only ever holds a few carries
    t = (uint32_t)s;
}
}  // namespace emul
#endif

// acc[0..7] += {a0,a1,a2,a3}*b on four consecutive 64-bit columns; carry-out added into acc[8].
HADES_DEV void cmad4(uint32_t (&acc)[9], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if !HADES_EMUL
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    const uint32_t a[4] = {a0, a1, a2, a3};
    emul::top(acc[8], emul::chain(acc, a, 0, b, 0));
#endif
}

// Same, preceded by  e0 += x  whose carry enters the chain.  `x` is the limb that falls off the
// other accumulator in the one-limb right shift; e0 and x share a limb position and acc[0] is
// the position right above it.
HADES_DEV void cmad4_shiftin(uint32_t (&acc)[9], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b,
                             uint32_t& e0, uint32_t x) {
#if !HADES_EMUL
    asm("add.cc.u32 %9, %9, %15;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(e0)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b), "r"(x));
#else
    uint64_t s = (uint64_t)e0 + x;
    e0 = (uint32_t)s;
    const uint32_t a[4] = {a0, a1, a2, a3};
    emul::top(acc[8], emul::chain(acc, a, 0, b, s >> 32));
#endif
}

// Montgomery step on the accumulator that owns limb 0.  With m = -acc[0] mod 2^32:
//   acc += m * (p0 + p2*2^64 + p4*2^128 + p6*2^192);  p0 = 1 so limb 0 becomes 0 and only its
//   carry (acc[0] != 0) matters.  3 products.  (The odd limbs p1,p3,p5,p7 go through cmad4.)
HADES_DEV void redc_even(uint32_t (&acc)[9], const MontQ& q) {
    // limb 0 (t0 + m = nz * 2^32) is not written: the caller drops it with the next shift
#if !HADES_EMUL
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %2;\n\t"
        "madc.lo.cc.u32 %3, %9, %11, %3;\n\t"
        "madc.hi.cc.u32 %4, %9, %11, %4;\n\t"
        "madc.lo.cc.u32 %5, %9, %12, %5;\n\t"
        "madc.hi.cc.u32 %6, %9, %12, %6;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
        : "r"(q.nz), "r"(q.m), "r"(modp(2)), "r"(modp(4)), "r"(modp(6)));
#else
    HADES_ASSERT((uint32_t)(acc[0] + q.m) == 0 && q.nz >= (acc[0] != 0 ? 1u : 0u) && q.nz <= 2u);
    uint64_t s = (uint64_t)acc[1] + q.nz;
    acc[1] = (uint32_t)s;
    const uint32_t a[4] = {0, p_limb(2), p_limb(4), p_limb(6)};
    emul::top(acc[8], emul::chain(acc, a, 1, q.m, s >> 32));
#endif
}

// Montgomery step on the accumulator that starts one limb higher ("odd": limb k = position k+1).
// The two low limbs of p are 2^64 - 2^32 + 1, so with t0 = even[0], m = -t0, nz = (t0 != 0):
//   t0 + m*(p0 + p1*2^32) = nz*2^32 + t0*2^32 + (m - nz)*2^64
// i.e. m*p[1] needs no product: position 1 gets +t0 (the +nz is the carry of even[0] + m, already
// added by redc_even) and position 2 gets +(m - nz).  3 products (p3, p5, p7) instead of 4.
HADES_DEV void redc_odd(uint32_t (&acc)[9], const MontQ& q) {
    const uint32_t t0 = q.t0, mm = q.mm, m = q.m;
#if !HADES_EMUL
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "madc.lo.cc.u32 %2, %11, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %11, %14, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
        : "r"(t0), "r"(mm), "r"(m), "r"(modp(3)), "r"(modp(5)), "r"(modp(7)));
#else
    uint64_t s = (uint64_t)acc[0] + t0;
    acc[0] = (uint32_t)s;
    s = (uint64_t)acc[1] + mm + (s >> 32);
    acc[1] = (uint32_t)s;
    const uint32_t a[4] = {0, p_limb(3), p_limb(5), p_limb(7)};
    emul::top(acc[8], emul::chain(acc, a, 1, m, s >> 32));
#endif
}

// r[0..8] = even + (odd << 32) + x   (x at limb 0); the discarded limb above r[8] must be zero.
HADES_DEV void merge_even_odd(uint32_t (&r)[9], const uint32_t (&e)[9], const uint32_t (&o)[9], uint32_t x) {
#if !HADES_EMUL
    asm("add.cc.u32 %0, %9, %26;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, %17, %25;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8])
        : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
          "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(x));
#else
    uint64_t s = (uint64_t)e[0] + x;
    r[0] = (uint32_t)s;
    for (int k = 1; k < 9; k++) {
        s = (uint64_t)e[k] + o[k - 1] + (s >> 32);
        r[k] = (uint32_t)s;
    }
    HADES_ASSERT((s >> 32) == 0 && o[8] == 0);
#endif
}

// (a[0..7], a8 as limb 8) - (p << S) over 9 limbs; difference (low 8 limbs) in d, limb 8 of the
// difference in d8; returns an all-ones MASK if the value was below p << S (final borrow), else 0.
template <int S>
HADES_DEV uint32_t sub_p_shl(uint32_t (&d)[8], uint32_t& d8, const uint32_t (&a)[8], uint32_t a8) {
    uint32_t below;
#if !HADES_EMUL
    uint32_t t8, t9;
    asm("sub.cc.u32 %0, %10, %19;\n\t"
        "subc.cc.u32 %1, %11, %20;\n\t"
        "subc.cc.u32 %2, %12, %21;\n\t"
        "subc.cc.u32 %3, %13, %22;\n\t"
        "subc.cc.u32 %4, %14, %23;\n\t"
        "subc.cc.u32 %5, %15, %24;\n\t"
        "subc.cc.u32 %6, %16, %25;\n\t"
        "subc.cc.u32 %7, %17, %26;\n\t"
        "subc.cc.u32 %8, %18, %27;\n\t"
        "subc.u32 %9, 0, 0;"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
          "=r"(t8), "=r"(t9)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a8),
          "r"(p_shl_limb(S, 0)), "r"(p_shl_limb(S, 1)), "r"(p_shl_limb(S, 2)), "r"(p_shl_limb(S, 3)),
          "r"(p_shl_limb(S, 4)), "r"(p_shl_limb(S, 5)), "r"(p_shl_limb(S, 6)), "r"(p_shl_limb(S, 7)),
          "r"(p_shl_limb(S, 8)));
    d8 = t8;
    below = t9;  // 0 - 0 - borrow: 0 or 0xffffffff
#else
    uint64_t bw = 0;
    for (int k = 0; k < 8; k++) {
        uint64_t t = (uint64_t)a[k] - p_shl_limb(S, k) - bw;
        d[k] = (uint32_t)t;
        bw = (t >> 63) & 1;
    }
    uint64_t t = (uint64_t)a8 - p_shl_limb(S, 8) - bw;
    d8 = (uint32_t)t;
    below = ((t >> 63) & 1) ? 0xffffffffu : 0u;
#endif
    return below;
}

// 8-limb version for values known to be below 2^256 and S = 0 (p itself): the 9th limb is skipped
HADES_DEV uint32_t sub_p8(uint32_t (&d)[8], const uint32_t (&a)[8]) {
    uint32_t below;
#if !HADES_EMUL
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(below)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(p_limb(0)), "r"(p_limb(1)), "r"(p_limb(2)), "r"(p_limb(3)), "r"(p_limb(4)), "r"(p_limb(5)), "r"(p_limb(6)),
          "r"(p_limb(7)));
#else
    uint64_t bw = 0;
    for (int k = 0; k < 8; k++) {
        uint64_t t = (uint64_t)a[k] - p_limb(k) - bw;
        d[k] = (uint32_t)t;
        bw = (t >> 63) & 1;
    }
    below = bw ? 0xffffffffu : 0u;
#endif
    return below;
}

// mask ? a : b per bit (one LOP3): the conditional subtractions select with the borrow MASK, which saves
// turning the borrow into a predicate (LOP3 + ISETP per subtraction)
HADES_DEV uint32_t bitsel(uint32_t mask, uint32_t a, uint32_t b) { return b ^ ((a ^ b) & mask); }

// r = a + b over 8 limbs, returns the carry-out limb (0/1).
HADES_DEV uint32_t add8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t carry;
#if !HADES_EMUL
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(carry)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) {
        s = (uint64_t)a[k] + b[k] + (s >> 32);
        r[k] = (uint32_t)s;
    }
    carry = (uint32_t)(s >> 32);
#endif
    return carry;
}

// r = a + b over 8 limbs for sums known to stay below 2^256 (no carry-out instruction)
HADES_DEV void add8_nc(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#if !HADES_EMUL
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint32_t c = add8(r, a, b);
    HADES_ASSERT(c == 0);
    (void)c;
#endif
}

// r = a + b + cin (cin = 0/1) over 8 limbs, returns the carry-out limb (0/1).
HADES_DEV uint32_t add8_cin(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], uint32_t cin) {
    uint32_t carry;
#if !HADES_EMUL
    asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"  // carry flag := cin
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=&r"(carry)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
#else
    uint64_t s = (uint64_t)cin << 32;
    for (int k = 0; k < 8; k++) {
        s = (uint64_t)a[k] + b[k] + (s >> 32);
        r[k] = (uint32_t)s;
    }
    carry = (uint32_t)(s >> 32);
#endif
    return carry;
}

// ------------------------------------------------------------------------------------------------
// Reductions to the canonical range
// ------------------------------------------------------------------------------------------------
// (x, x8) -> (x, x8) - (p << S) if that is non-negative.
template <int S>
HADES_DEV void cond_sub_p_shl(uint32_t (&x)[8], uint32_t& x8) {
    uint32_t d[8], d8;
    uint32_t below = sub_p_shl<S>(d, d8, x, x8);
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = bitsel(below, x[k], d[k]);
    x8 = bitsel(below, x8, d8);
}
// x -> x - p if that is non-negative, for x < 2^256
HADES_DEV void cond_sub_p8(uint32_t (&x)[8]) {
    uint32_t d[8];
    uint32_t below = sub_p8(d, x);
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = bitsel(below, x[k], d[k]);
}

// r (9 limbs, value < 2^(LOG2+1) * p... precisely value < 2p << LOG2) -> canonical [0,p).
// LOG2 = 0: value < 2p; 1: < 4p; 2: < 8p; 3: < 16p; 4: < 32p (all fit the 9 limbs).
template <int LOG2>
HADES_DEV void canon(Fr& out, const uint32_t (&r)[9]) {
    static_assert(LOG2 >= 0 && LOG2 <= 4, "bound out of range");
    uint32_t x[8], x8 = r[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = r[k];
    if constexpr (LOG2 >= 4) cond_sub_p_shl<4>(x, x8);
    if constexpr (LOG2 >= 3) cond_sub_p_shl<3>(x, x8);
    if constexpr (LOG2 >= 2) cond_sub_p_shl<2>(x, x8);
    if constexpr (LOG2 >= 1) {
        cond_sub_p_shl<1>(x, x8);
        HADES_ASSERT(x8 == 0);  // < 2p < 2^256: the last subtraction runs on 8 limbs
        cond_sub_p8(x);
    } else {
        cond_sub_p_shl<0>(x, x8);
        HADES_ASSERT(x8 == 0);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out.l[k] = x[k];
}

// smallest LOG2 such that a value < bound_p * p is below 2p << LOG2
HADES_DEV constexpr int canon_log2_for(int bound_p) {
    return bound_p <= 2 ? 0 : bound_p <= 4 ? 1 : bound_p <= 8 ? 2 : bound_p <= 16 ? 3 : 4;
}

// out = a + b mod p, inputs canonical (scalar.rs:28 `*w += c`)
HADES_DEV void fr_add(Fr& out, const Fr& a, const Fr& b) {
    uint32_t s[8];
    add8_nc(s, a.l, b.l);  // a + b < 2p < 2^256
    cond_sub_p8(s);
#pragma unroll
    for (int k = 0; k < 8; k++) out.l[k] = s[k];
}

// ------------------------------------------------------------------------------------------------
// N-term Montgomery dot product:  r = (sum_j A_j * B_j) / 2^256  (mod p, NOT yet canonical).
// `vec(j,k)` yields limb k of the operand walked by the inner chains, `sca(j,i)` limb i of the
// operand consumed one limb per outer step.  Result: 9 limbs, value < p * (1 + sum_j A_j*B_j/(p*R)).
// Even/odd bookkeeping per outer step i (all indices compile-time, so the "shift" is renaming):
//   E holds limb positions 0..8 of the running sum, O positions 1..9, X one pending limb at 0.
// ------------------------------------------------------------------------------------------------
template <int N, bool kFirst, class Vec, class Sca>
HADES_DEV void dot_step(uint32_t (&E)[9], uint32_t (&O)[9], uint32_t x, int i, Vec vec, Sca sca) {
    // odd limbs of every term -> O (positions 1..8); the first chain also folds the pending limb in
    if constexpr (kFirst) {
        cmad4(O, vec(0, 1), vec(0, 3), vec(0, 5), vec(0, 7), sca(0, i));
    } else {
        cmad4_shiftin(O, vec(0, 1), vec(0, 3), vec(0, 5), vec(0, 7), sca(0, i), E[0], x);
    }
#pragma unroll
    for (int j = 1; j < N; j++) cmad4(O, vec(j, 1), vec(j, 3), vec(j, 5), vec(j, 7), sca(j, i));
    // even limbs -> E (positions 0..7)
#pragma unroll
    for (int j = 0; j < N; j++) cmad4(E, vec(j, 0), vec(j, 2), vec(j, 4), vec(j, 6), sca(j, i));
    // Montgomery step: add m*p so that position 0 clears
    const MontQ q = mont_quotient(E[0]);
    redc_odd(O, q);
    redc_even(E, q);
}

// With kInject the 16-limb integer `t` is added to the sum before the reduction (STEPS = 8 only):
//     r = (t + sum_j A_j * B_j) / 2^256.
// Its low limbs are the initial accumulator contents and each higher limb enters the limb that the
// one-limb shift has just vacated, so the addition costs no instruction at all (cf. redc16).
template <int N, int STEPS, bool kInject, class Vec, class Sca>
HADES_DEV void dot_mont_core(uint32_t (&r)[9], Vec vec, Sca sca, const uint32_t* t) {
    static_assert(STEPS == 2 || STEPS == 4 || STEPS == 8, "even/odd bookkeeping needs an even step count");
    static_assert(!kInject || STEPS == 8, "the addend is a full 512-bit integer");
    uint32_t A[9], B[9];
#pragma unroll
    for (int k = 0; k < 9; k++) A[k] = B[k] = 0;
    if constexpr (kInject) {
#pragma unroll
        for (int k = 0; k < 8; k++) A[k] = t[k];
        B[7] = t[8];  // odd accumulator limb k sits at position k + 1
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < STEPS; i += 2) {
        // step i: E = A, O = B
        if (i == 0) dot_step<N, true>(A, B, 0u, i, vec, sca);
        else dot_step<N, false>(A, B, x, i, vec, sca);
        // shift one limb: new E = B (positions 1..9 -> 0..8); new O[k] = A[k+2]; pending = A[1]
        x = A[1];
#pragma unroll
        for (int k = 0; k < 7; k++) A[k] = A[k + 2];
        A[7] = kInject ? t[i + 9] : 0u;
        A[8] = 0;
        // step i+1: E = B, O = A
        dot_step<N, false>(B, A, x, i + 1, vec, sca);
        x = B[1];
#pragma unroll
        for (int k = 0; k < 7; k++) B[k] = B[k + 2];
        B[7] = (kInject && i + 10 < 16) ? t[i + 10] : 0u;
        B[8] = 0;
    }
    // after the steps: E = A (positions 0..8), O = B (positions 1..9), pending x at position 0
    merge_even_odd(r, A, B, x);
}

template <int N, int STEPS, class Vec, class Sca>
HADES_DEV void dot_mont_steps(uint32_t (&r)[9], Vec vec, Sca sca) {
    dot_mont_core<N, STEPS, false>(r, vec, sca, nullptr);
}

// r = (t + sum_j A_j * B_j) / 2^256 for a 512-bit integer t (one reduction for the sum and the addend)
template <int N, class Vec, class Sca>
HADES_DEV void dot_mont_plus(uint32_t (&r)[9], Vec vec, Sca sca, const uint32_t (&t)[16]) {
    dot_mont_core<N, 8, true>(r, vec, sca, t);
}

template <int N, class Vec, class Sca>
HADES_DEV void dot_mont(uint32_t (&r)[9], Vec vec, Sca sca) {
    dot_mont_steps<N, 8>(r, vec, sca);
}

// r = c * y / 2^256 (mod p, < 5p for K = 4, < 3p for K = 2) for a constant c given as its K short-reduction
// versions: xk(j, limb) = limb of X_j = c * 2^(256*(j+1)/K - 256) mod p, j = 0..K-1.
template <int K, class XK>
HADES_DEV void mul_const_short(uint32_t (&r)[9], XK xk, const Fr& y) {
    constexpr int kSteps = 8 / K;
    dot_mont_steps<K, kSteps>(
        r, [&](int j, int limb) { return xk(j, limb); }, [&](int j, int i) { return y.l[j * kSteps + i]; });
}

// ------------------------------------------------------------------------------------------------
// Dedicated Montgomery squaring: 28 off-diagonal + 8 diagonal + 48 reduction = 84 products (a general
// product costs 64 + 48 = 112).
// ------------------------------------------------------------------------------------------------
// Short chains for the triangle a_i*a_j (i<j): L links on consecutive 64-bit columns, carry-out added
// to the limb right above (`land`).  Rows are processed in ascending i, which guarantees that a landing
// limb only ever holds a few carries when it is hit (never a product limb), so no ripple is possible.
HADES_DEV void cmad1(uint32_t& l0, uint32_t& h0, uint32_t& land, uint32_t a0, uint32_t b) {
#if !HADES_EMUL
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(l0), "+r"(h0), "+r"(land)
        : "r"(a0), "r"(b));
#else
    uint32_t acc[2] = {l0, h0};
    const uint32_t a[4] = {0, 0, 0, a0};
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, acc[0], acc[1]};
    uint64_t c = emul::chain(w, a, 3, b, 0);
    l0 = w[6]; h0 = w[7];
    emul::top(land, c);
#endif
}
HADES_DEV void cmad2(uint32_t& l0, uint32_t& h0, uint32_t& l1, uint32_t& h1, uint32_t& land, uint32_t a0, uint32_t a1,
                     uint32_t b) {
#if !HADES_EMUL
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(l0), "+r"(h0), "+r"(l1), "+r"(h1), "+r"(land)
        : "r"(a0), "r"(a1), "r"(b));
#else
    const uint32_t a[4] = {0, 0, a0, a1};
    uint32_t w[8] = {0, 0, 0, 0, l0, h0, l1, h1};
    uint64_t c = emul::chain(w, a, 2, b, 0);
    l0 = w[4]; h0 = w[5]; l1 = w[6]; h1 = w[7];
    emul::top(land, c);
#endif
}
HADES_DEV void cmad3(uint32_t& l0, uint32_t& h0, uint32_t& l1, uint32_t& h1, uint32_t& l2, uint32_t& h2, uint32_t& land,
                     uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
#if !HADES_EMUL
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(l0), "+r"(h0), "+r"(l1), "+r"(h1), "+r"(l2), "+r"(h2), "+r"(land)
        : "r"(a0), "r"(a1), "r"(a2), "r"(b));
#else
    const uint32_t a[4] = {0, a0, a1, a2};
    uint32_t w[8] = {0, 0, l0, h0, l1, h1, l2, h2};
    uint64_t c = emul::chain(w, a, 1, b, 0);
    l0 = w[2]; h0 = w[3]; l1 = w[4]; h1 = w[5]; l2 = w[6]; h2 = w[7];
    emul::top(land, c);
#endif
}
HADES_DEV void cmad4r(uint32_t& l0, uint32_t& h0, uint32_t& l1, uint32_t& h1, uint32_t& l2, uint32_t& h2, uint32_t& l3,
                      uint32_t& h3, uint32_t& land, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    uint32_t acc[9] = {l0, h0, l1, h1, l2, h2, l3, h3, land};
    cmad4(acc, a0, a1, a2, a3, b);
    l0 = acc[0]; h0 = acc[1]; l1 = acc[2]; h1 = acc[3]; l2 = acc[4]; h2 = acc[5]; l3 = acc[6]; h3 = acc[7]; land = acc[8];
}

// t[0..15] = a * b as a plain 512-bit integer (64 products, no reduction).  Two accumulators of 64-bit columns,
// EA aligned at even limb positions and OA at odd ones (OA[k] = position k + 1); rows in ascending i, so a
// chain's carry-out always lands in a limb that holds nothing but carries.
HADES_DEV void mul_wide(uint32_t (&t)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t EA[17], OA[16];
#pragma unroll
    for (int k = 0; k < 17; k++) EA[k] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) OA[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        // b_i, i even: even limbs of a -> positions i + 2k (EA), odd limbs -> positions i + 1 + 2k (OA index i + 2k)
        cmad4r(EA[i], EA[i + 1], EA[i + 2], EA[i + 3], EA[i + 4], EA[i + 5], EA[i + 6], EA[i + 7], EA[i + 8], a[0], a[2], a[4], a[6], b[i]);
        cmad4r(OA[i], OA[i + 1], OA[i + 2], OA[i + 3], OA[i + 4], OA[i + 5], OA[i + 6], OA[i + 7], OA[i + 8], a[1], a[3], a[5], a[7], b[i]);
        // b_{i+1}: even limbs of a -> positions i + 1 + 2k (OA index i + 2k), odd limbs -> positions i + 2 + 2k (EA)
        cmad4r(OA[i], OA[i + 1], OA[i + 2], OA[i + 3], OA[i + 4], OA[i + 5], OA[i + 6], OA[i + 7], OA[i + 8], a[0], a[2], a[4], a[6], b[i + 1]);
        cmad4r(EA[i + 2], EA[i + 3], EA[i + 4], EA[i + 5], EA[i + 6], EA[i + 7], EA[i + 8], EA[i + 9], EA[i + 10], a[1], a[3], a[5], a[7], b[i + 1]);
    }
    uint32_t lo[8], hi[8], e_lo[8], e_hi[8], o_lo[8], o_hi[8];
    o_lo[0] = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) e_lo[k] = EA[k];
#pragma unroll
    for (int k = 1; k < 8; k++) o_lo[k] = OA[k - 1];
#pragma unroll
    for (int k = 0; k < 8; k++) { e_hi[k] = EA[8 + k]; o_hi[k] = OA[7 + k]; }
    uint32_t c = add8(lo, e_lo, o_lo);
    uint32_t c2 = add8_cin(hi, e_hi, o_hi, c);
    HADES_ASSERT(c2 == 0 && EA[16] == 0 && OA[15] == 0);  // a * b < 2^512
    (void)c2;
#pragma unroll
    for (int k = 0; k < 8; k++) { t[k] = lo[k]; t[8 + k] = hi[k]; }
}

// t[0..15] = 2*t + sum_i a_i^2 * 2^(64 i)   (t < 2^511 on entry; the result a^2-part fits 512 bits)
HADES_DEV void double_add_diag(uint32_t (&t)[16], const uint32_t (&a)[8]) {
#if !HADES_EMUL
    uint32_t d[16];
    d[0] = t[0] << 1;
#pragma unroll
    for (int k = 1; k < 16; k++) d[k] = __funnelshift_l(t[k - 1], t[k], 1);
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]), "+r"(d[8]),
          "+r"(d[9]), "+r"(d[10]), "+r"(d[11]), "+r"(d[12]), "+r"(d[13]), "+r"(d[14]), "+r"(d[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = d[k];
#else
    HADES_ASSERT((t[15] >> 31) == 0);
    uint32_t d[16];
    d[0] = t[0] << 1;
    for (int k = 1; k < 16; k++) d[k] = (t[k] << 1) | (t[k - 1] >> 31);
    uint64_t carry = 0;
    for (int i = 0; i < 8; i++) {
        unsigned __int128 v = (unsigned __int128)a[i] * a[i] + (((uint64_t)d[2 * i + 1] << 32) | d[2 * i]) + carry;
        d[2 * i] = (uint32_t)v;
        d[2 * i + 1] = (uint32_t)(v >> 32);
        carry = (uint64_t)(v >> 64);
    }
    HADES_ASSERT(carry == 0);
    for (int k = 0; k < 16; k++) t[k] = d[k];
#endif
}

// e0 += x, returns the carry (0/1)
HADES_DEV uint32_t add_carry_out(uint32_t& e0, uint32_t x) {
    uint32_t c;
#if !HADES_EMUL
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(e0), "=r"(c) : "r"(x));
#else
    uint64_t s = (uint64_t)e0 + x;
    e0 = (uint32_t)s;
    c = (uint32_t)(s >> 32);
#endif
    return c;
}

// r = t / 2^256 mod p (9 limbs, < t/2^256 + p) for a merged 512-bit t: the reduce-only version of
// dot_mont -- same even/odd bookkeeping, the upper limbs of t are injected one per shift.
HADES_DEV void redc16(uint32_t (&r)[9], const uint32_t (&t)[16]) {
    uint32_t A[9], B[9];
#pragma unroll
    for (int k = 0; k < 8; k++) { A[k] = t[k]; B[k] = 0; }
    A[8] = 0; B[8] = 0;
    B[7] = t[8];  // odd accumulator limb k sits at position k + 1
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        {  // step i: E = A, O = B
            uint32_t cx = (i == 0) ? 0u : add_carry_out(A[0], x);
            MontQ q = mont_quotient(A[0]);
            redc_odd(B, q);
            q.nz += cx;  // both carries enter limb 1
            redc_even(A, q);
            x = A[1];
#pragma unroll
            for (int k = 0; k < 7; k++) A[k] = A[k + 2];
            A[7] = t[i + 9];
            A[8] = 0;
        }
        {  // step i+1: E = B, O = A
            uint32_t cx = add_carry_out(B[0], x);
            MontQ q = mont_quotient(B[0]);
            redc_odd(A, q);
            q.nz += cx;
            redc_even(B, q);
            x = B[1];
#pragma unroll
            for (int k = 0; k < 7; k++) B[k] = B[k + 2];
            B[7] = (i + 10 < 16) ? t[i + 10] : 0u;
            B[8] = 0;
        }
    }
    merge_even_odd(r, A, B, x);
}

// r = a^2 / 2^256 mod p, 9 limbs, value < a^2/2^256 + p
HADES_DEV void sqr_mont(uint32_t (&r)[9], const uint32_t (&a)[8]) {
    uint32_t E[17], O[16];  // E[k] = position k, O[k] = position k + 1
#pragma unroll
    for (int k = 0; k < 17; k++) E[k] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) O[k] = 0;
    // row 0: a0 * a[1..7]
    cmad4r(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], O[8], a[1], a[3], a[5], a[7], a[0]);
    cmad3(E[2], E[3], E[4], E[5], E[6], E[7], E[8], a[2], a[4], a[6], a[0]);
    // row 1: a1 * a[2..7]
    cmad3(E[4], E[5], E[6], E[7], E[8], E[9], E[10], a[3], a[5], a[7], a[1]);
    cmad3(O[2], O[3], O[4], O[5], O[6], O[7], O[8], a[2], a[4], a[6], a[1]);
    // row 2: a2 * a[3..7]
    cmad3(O[4], O[5], O[6], O[7], O[8], O[9], O[10], a[3], a[5], a[7], a[2]);
    cmad2(E[6], E[7], E[8], E[9], E[10], a[4], a[6], a[2]);
    // row 3: a3 * a[4..7]
    cmad2(E[8], E[9], E[10], E[11], E[12], a[5], a[7], a[3]);
    cmad2(O[6], O[7], O[8], O[9], O[10], a[4], a[6], a[3]);
    // row 4: a4 * a[5..7]
    cmad2(O[8], O[9], O[10], O[11], O[12], a[5], a[7], a[4]);
    cmad1(E[10], E[11], E[12], a[6], a[4]);
    // row 5: a5 * a[6..7]
    cmad1(E[12], E[13], E[14], a[7], a[5]);
    cmad1(O[10], O[11], O[12], a[6], a[5]);
    // row 6: a6 * a7
    cmad1(O[12], O[13], O[14], a[7], a[6]);
    // merge: t = E + (O << 32), 16 limbs (the triangle is < 2^511)
    uint32_t t[16];
    {
        uint32_t lo[8], hi[8], e_lo[8], e_hi[8], o_lo[8], o_hi[8];
        e_lo[0] = E[0];
        o_lo[0] = 0;
#pragma unroll
        for (int k = 1; k < 8; k++) { e_lo[k] = E[k]; o_lo[k] = O[k - 1]; }
        e_lo[0] = E[0];
#pragma unroll
        for (int k = 0; k < 8; k++) { e_hi[k] = E[8 + k]; o_hi[k] = O[7 + k]; }
        uint32_t c = add8(lo, e_lo, o_lo);
        uint32_t c2 = add8_cin(hi, e_hi, o_hi, c);
        HADES_ASSERT(c2 == 0 && E[16] == 0 && O[15] == 0);  // the triangle is < 2^511
        (void)c2;
#pragma unroll
        for (int k = 0; k < 8; k++) { t[k] = lo[k]; t[8 + k] = hi[k]; }
    }
    double_add_diag(t, a);
    redc16(r, t);
}

// out = a*b/R mod p, canonical.  a,b canonical (or any a,b with a*b < p*R).
HADES_DEV void fr_mul(Fr& out, const Fr& a, const Fr& b) {
    uint32_t r[9];
    dot_mont<1>(r, [&](int, int k) { return a.l[k]; }, [&](int, int i) { return b.l[i]; });
    canon<0>(out, r);
}

// non-canonical product for internal chains: result < p*(1 + a*b/(p*R)) < 2^256 when a,b < 2p-ish
HADES_DEV void fr_mul_lazy(Fr& out, const Fr& a, const Fr& b) {
    uint32_t r[9];
    dot_mont<1>(r, [&](int, int k) { return a.l[k]; }, [&](int, int i) { return b.l[i]; });
    HADES_ASSERT(r[8] == 0);
#pragma unroll
    for (int k = 0; k < 8; k++) out.l[k] = r[k];
}

// x -> x^5 (scalar.rs:32-34 `value.square().square() * value`), canonical in, canonical out.
// Bounds with x < p (p/R = 0.4528): x^2 < 1.453p, x^4 < 1.956p, x^5 < 1.886p -- all below 2^256,
// so only the last product needs the conditional subtraction.
HADES_DEV void fr_sqr_lazy(Fr& out, const Fr& a) {
    uint32_t r[9];
    sqr_mont(r, a.l);
    HADES_ASSERT(r[8] == 0);
#pragma unroll
    for (int k = 0; k < 8; k++) out.l[k] = r[k];
}

// x^4 by two trips through ONE squaring body (a real loop: the kernels are instruction-cache sensitive)
HADES_DEV void fr_pow4_lazy(Fr& x4, const Fr& x) {
    x4 = x;
#if !HADES_EMUL
#pragma unroll 1
#endif
    for (int i = 0; i < 2; i++) fr_sqr_lazy(x4, x4);
}

HADES_DEV void fr_sbox(Fr& x) {
    Fr x4;
    fr_pow4_lazy(x4, x);
    fr_mul(x, x4, x);
}

}  // namespace hades
