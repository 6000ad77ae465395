// One level of Karatsuba for the constant-times-state dot products of the canonical-form partial rounds
// (hades.cuh partial_round_ccf; the reference's hot loop is src/strategies/scalar.rs:36-49).
//
//   T = Y + sum_j C_j * W_j        (C_j: table constants, W_j: state words, Y: the 512-bit S-box product)
//
// With 128-bit halves C = Cl + Ch*2^128, W = Wl + Wh*2^128:
//   C*W = Cl*Wl + (Cl*Wh + Ch*Wl)*2^128 + Ch*Wh*2^256 ,   Cl*Wh + Ch*Wl = (Cl+Ch)*(Wl+Wh) - Cl*Wl - Ch*Wh
// so a term costs three 4x4-limb products (48 IMAD.WIDE) instead of 64, and because the recombination is linear it
// is done ONCE per dot product on the three accumulated sums
//   LL = sum Cl*Wl ,  HH = sum Ch*Wh ,  MM = sum (Cl+Ch)*(Wl+Wh)   =>   T = Y + LL + (MM - LL - HH)*2^128 + HH*2^256.
// The 129th bits: Cl+Ch = CS + kappa*2^128 (host-side, stored in the table: uniform), Wl+Wh = WS + gamma*2^128
// (per thread), (Cl+Ch)*(Wl+Wh) = CS*WS + (kappa*WS + gamma*CS)*2^128 + kappa*gamma*2^256: the two middle terms are
// conditional 4-limb additions (a uniform branch for kappa, predicated adds for gamma), no products.
// Y costs nothing: its halves are the initial contents of the LL and HH accumulators (injected limb by limb as the
// accumulator window moves up), and MM starts from Y_lo + Y_hi so that the subtraction cancels it.
// The 17-limb T is then reduced by redc17 (48 products): 4*48 + 48 = 240 products per 4-term dot product instead
// of 4*64 + 48 = 304.  T is the same integer as in dot_mont_plus, the Montgomery quotient is unique, so the 9-limb
// result is IDENTICAL to dot_mont_plus (checked limb for limb in the emulation build and on the device).
#pragma once
#include "fr.cuh"

namespace hades {

// acc[0..3] += {a0,a1}*b on two consecutive 64-bit columns; carry-out added into acc[4] (which only holds carries)
HADES_DEV void cmad2w(uint32_t (&acc)[5], uint32_t a0, uint32_t a1, uint32_t b) {
#if !HADES_EMUL
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
        : "r"(a0), "r"(a1), "r"(b));
#else
    uint64_t carry = 0;
    const uint32_t a[2] = {a0, a1};
    for (int k = 0; k < 2; k++) {
        unsigned __int128 t = (unsigned __int128)a[k] * b + (((uint64_t)acc[2 * k + 1] << 32) | acc[2 * k]) + carry;
        acc[2 * k] = (uint32_t)t;
        acc[2 * k + 1] = (uint32_t)(t >> 32);
        carry = (uint64_t)(t >> 64);
    }
    emul::top(acc[4], carry);
#endif
}
// same, preceded by  e0 += x  whose carry enters the chain (cf. cmad4_shiftin)
HADES_DEV void cmad2w_shiftin(uint32_t (&acc)[5], uint32_t a0, uint32_t a1, uint32_t b, uint32_t& e0, uint32_t x) {
#if !HADES_EMUL
    asm("add.cc.u32 %5, %5, %9;\n\t"
        "madc.lo.cc.u32 %0, %6, %8, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %8, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, %8, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(e0)
        : "r"(a0), "r"(a1), "r"(b), "r"(x));
#else
    uint64_t s = (uint64_t)e0 + x;
    e0 = (uint32_t)s;
    uint64_t carry = s >> 32;
    const uint32_t a[2] = {a0, a1};
    for (int k = 0; k < 2; k++) {
        unsigned __int128 t = (unsigned __int128)a[k] * b + (((uint64_t)acc[2 * k + 1] << 32) | acc[2 * k]) + carry;
        acc[2 * k] = (uint32_t)t;
        acc[2 * k + 1] = (uint32_t)(t >> 32);
        carry = (uint64_t)(t >> 64);
    }
    emul::top(acc[4], carry);
#endif
}

// r[0..N-1] = a + b (+ cin) over N limbs; the carry out of limb N-1 must be zero (asserted in the emulation build)
template <int N>
HADES_DEV void addn_nc(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N]);
template <>
HADES_DEV void addn_nc<9>(uint32_t (&r)[9], const uint32_t (&a)[9], const uint32_t (&b)[9]) {
#if !HADES_EMUL
    asm("add.cc.u32 %0, %9, %18;\n\t"
        "addc.cc.u32 %1, %10, %19;\n\t"
        "addc.cc.u32 %2, %11, %20;\n\t"
        "addc.cc.u32 %3, %12, %21;\n\t"
        "addc.cc.u32 %4, %13, %22;\n\t"
        "addc.cc.u32 %5, %14, %23;\n\t"
        "addc.cc.u32 %6, %15, %24;\n\t"
        "addc.cc.u32 %7, %16, %25;\n\t"
        "addc.u32 %8, %17, %26;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]));
#else
    uint64_t s = 0;
    for (int k = 0; k < 9; k++) {
        s = (uint64_t)a[k] + b[k] + (s >> 32);
        r[k] = (uint32_t)s;
    }
    HADES_ASSERT((s >> 32) == 0);
#endif
}
// r = a - b over 9 limbs, a >= b (asserted in the emulation build)
HADES_DEV void sub9_nb(uint32_t (&r)[9], const uint32_t (&a)[9], const uint32_t (&b)[9]) {
#if !HADES_EMUL
    asm("sub.cc.u32 %0, %9, %18;\n\t"
        "subc.cc.u32 %1, %10, %19;\n\t"
        "subc.cc.u32 %2, %11, %20;\n\t"
        "subc.cc.u32 %3, %12, %21;\n\t"
        "subc.cc.u32 %4, %13, %22;\n\t"
        "subc.cc.u32 %5, %14, %23;\n\t"
        "subc.cc.u32 %6, %15, %24;\n\t"
        "subc.cc.u32 %7, %16, %25;\n\t"
        "subc.u32 %8, %17, %26;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]));
#else
    uint64_t bw = 0;
    for (int k = 0; k < 9; k++) {
        uint64_t t = (uint64_t)a[k] - b[k] - bw;
        r[k] = (uint32_t)t;
        bw = (t >> 63) & 1;
    }
    HADES_ASSERT(bw == 0);
#endif
}

// ws = lo + hi over 4 limbs, returns the carry (0/1)
HADES_DEV uint32_t add4_carry(uint32_t (&ws)[4], uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t h0, uint32_t h1,
                              uint32_t h2, uint32_t h3) {
    uint32_t c;
#if !HADES_EMUL
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(ws[0]), "=r"(ws[1]), "=r"(ws[2]), "=r"(ws[3]), "=r"(c)
        : "r"(l0), "r"(l1), "r"(l2), "r"(l3), "r"(h0), "r"(h1), "r"(h2), "r"(h3));
#else
    const uint32_t l[4] = {l0, l1, l2, l3}, h[4] = {h0, h1, h2, h3};
    uint64_t s = 0;
    for (int k = 0; k < 4; k++) {
        s = (uint64_t)l[k] + h[k] + (s >> 32);
        ws[k] = (uint32_t)s;
    }
    c = (uint32_t)(s >> 32);
#endif
    return c;
}

// if (flag) K += {v0..v3} (4 limbs into a 5-limb value; limb 4 only collects carries).  `flag` differs per
// thread: predicated adds, no branch.
HADES_DEV void cond_add4(uint32_t (&K)[5], uint32_t flag, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
#if !HADES_EMUL
    asm("{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.u32 q, %5, 0;\n\t"
        "@q add.cc.u32 %0, %0, %6;\n\t"
        "@q addc.cc.u32 %1, %1, %7;\n\t"
        "@q addc.cc.u32 %2, %2, %8;\n\t"
        "@q addc.cc.u32 %3, %3, %9;\n\t"
        "@q addc.u32 %4, %4, 0;\n\t"
        "}"
        : "+r"(K[0]), "+r"(K[1]), "+r"(K[2]), "+r"(K[3]), "+r"(K[4])
        : "r"(flag), "r"(v0), "r"(v1), "r"(v2), "r"(v3));
#else
    if (flag) {
        const uint32_t v[4] = {v0, v1, v2, v3};
        uint64_t s = 0;
        for (int k = 0; k < 4; k++) {
            s = (uint64_t)K[k] + v[k] + (s >> 32);
            K[k] = (uint32_t)s;
        }
        K[4] += (uint32_t)(s >> 32);
    }
#endif
}
// K += {v0..v3} unconditionally (inside a uniform branch)
HADES_DEV void add4_into5(uint32_t (&K)[5], uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
#if !HADES_EMUL
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(K[0]), "+r"(K[1]), "+r"(K[2]), "+r"(K[3]), "+r"(K[4])
        : "r"(v0), "r"(v1), "r"(v2), "r"(v3));
#else
    cond_add4(K, 1u, v0, v1, v2, v3);
#endif
}

// out[0..8] = t[0..8] + sum_j vec_j * sca_j  for 4-limb operands (vec(j,k), sca(j,i), k,i < 4), as a plain integer.
// Same even/odd window as dot_mont_core, but a step RETIRES its lowest limb instead of cancelling it.
template <int N, bool kInject, class Vec, class Sca>
HADES_DEV void wide4_acc(uint32_t (&out)[9], Vec vec, Sca sca, const uint32_t* t) {
    uint32_t A[5], B[5];
#pragma unroll
    for (int k = 0; k < 5; k++) A[k] = B[k] = 0;
    if constexpr (kInject) {
#pragma unroll
        for (int k = 0; k < 4; k++) A[k] = t[k];
        B[3] = t[4];  // odd accumulator limb k sits at position k + 1
    }
    uint32_t x = 0;
    auto step = [&](uint32_t (&E)[5], uint32_t (&O)[5], int i, bool first) {
        if (first) cmad2w(O, vec(0, 1), vec(0, 3), sca(0, i));
        else cmad2w_shiftin(O, vec(0, 1), vec(0, 3), sca(0, i), E[0], x);
#pragma unroll
        for (int j = 1; j < N; j++) cmad2w(O, vec(j, 1), vec(j, 3), sca(j, i));
#pragma unroll
        for (int j = 0; j < N; j++) cmad2w(E, vec(j, 0), vec(j, 2), sca(j, i));
    };
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        step(A, B, i, i == 0);
        out[i] = A[0];
        x = A[1];
        A[0] = A[2]; A[1] = A[3]; A[2] = A[4];
        A[3] = kInject ? t[i + 5] : 0u;
        A[4] = 0;
        step(B, A, i + 1, false);
        out[i + 1] = B[0];
        x = B[1];
        B[0] = B[2]; B[1] = B[3]; B[2] = B[4];
        B[3] = kInject ? t[i + 6] : 0u;
        B[4] = 0;
    }
    // E = A (positions 4..8), O = B (positions 5..9, limb 4 must stay zero), pending x at position 4
#if !HADES_EMUL
    asm("add.cc.u32 %0, %5, %14;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, %9, %13;"
        : "=r"(out[4]), "=r"(out[5]), "=r"(out[6]), "=r"(out[7]), "=r"(out[8])
        : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(B[0]), "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(x));
#else
    uint64_t s = (uint64_t)A[0] + x;
    out[4] = (uint32_t)s;
    for (int k = 1; k < 5; k++) {
        s = (uint64_t)A[k] + B[k - 1] + (s >> 32);
        out[4 + k] = (uint32_t)s;
    }
    HADES_ASSERT((s >> 32) == 0 && B[4] == 0);
#endif
}

// T[0..16] = L + D*2^128 + H*2^256 for 9-limb L, D, H (the sum must fit 17 limbs)
HADES_DEV void kara_assemble(uint32_t (&T)[17], const uint32_t (&L)[9], const uint32_t (&D)[9], const uint32_t (&H)[9]) {
    uint32_t X[9];  // H + L[8]
#if !HADES_EMUL
    asm("add.cc.u32 %0, %9, %18;\n\t"
        "addc.cc.u32 %1, %10, 0;\n\t"
        "addc.cc.u32 %2, %11, 0;\n\t"
        "addc.cc.u32 %3, %12, 0;\n\t"
        "addc.cc.u32 %4, %13, 0;\n\t"
        "addc.cc.u32 %5, %14, 0;\n\t"
        "addc.cc.u32 %6, %15, 0;\n\t"
        "addc.cc.u32 %7, %16, 0;\n\t"
        "addc.u32 %8, %17, 0;"
        : "=r"(X[0]), "=r"(X[1]), "=r"(X[2]), "=r"(X[3]), "=r"(X[4]), "=r"(X[5]), "=r"(X[6]), "=r"(X[7]), "=r"(X[8])
        : "r"(H[0]), "r"(H[1]), "r"(H[2]), "r"(H[3]), "r"(H[4]), "r"(H[5]), "r"(H[6]), "r"(H[7]), "r"(H[8]), "r"(L[8]));
    asm("add.cc.u32 %0, %13, %26;\n\t"
        "addc.cc.u32 %1, %14, %27;\n\t"
        "addc.cc.u32 %2, %15, %28;\n\t"
        "addc.cc.u32 %3, %16, %29;\n\t"
        "addc.cc.u32 %4, %17, %30;\n\t"
        "addc.cc.u32 %5, %18, %31;\n\t"
        "addc.cc.u32 %6, %19, %32;\n\t"
        "addc.cc.u32 %7, %20, %33;\n\t"
        "addc.cc.u32 %8, %21, %34;\n\t"
        "addc.cc.u32 %9, %22, 0;\n\t"
        "addc.cc.u32 %10, %23, 0;\n\t"
        "addc.cc.u32 %11, %24, 0;\n\t"
        "addc.u32 %12, %25, 0;"
        : "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]),
          "=r"(T[13]), "=r"(T[14]), "=r"(T[15]), "=r"(T[16])
        : "r"(L[4]), "r"(L[5]), "r"(L[6]), "r"(L[7]), "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]),
          "r"(X[6]), "r"(X[7]), "r"(X[8]), "r"(D[0]), "r"(D[1]), "r"(D[2]), "r"(D[3]), "r"(D[4]), "r"(D[5]), "r"(D[6]),
          "r"(D[7]), "r"(D[8]));
#else
    uint64_t s = (uint64_t)H[0] + L[8];
    X[0] = (uint32_t)s;
    for (int k = 1; k < 9; k++) {
        s = (uint64_t)H[k] + (s >> 32);
        X[k] = (uint32_t)s;
    }
    HADES_ASSERT((s >> 32) == 0);
    s = 0;
    for (int k = 0; k < 13; k++) {
        const uint32_t a = k < 4 ? L[4 + k] : X[k - 4];
        const uint32_t d = k < 9 ? D[k] : 0u;
        s = (uint64_t)a + d + (s >> 32);
        T[4 + k] = (uint32_t)s;
    }
    HADES_ASSERT((s >> 32) == 0);
#endif
#pragma unroll
    for (int k = 0; k < 4; k++) T[k] = L[k];
}

// r = t / 2^256 mod p for a 17-limb t (9 limbs, < t/2^256 + p): redc16 with one more injected limb
HADES_DEV void redc17(uint32_t (&r)[9], const uint32_t (&t)[17]) {
    uint32_t A[9], B[9];
#pragma unroll
    for (int k = 0; k < 8; k++) { A[k] = t[k]; B[k] = 0; }
    A[8] = 0; B[8] = 0;
    B[7] = t[8];
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        {
            uint32_t cx = (i == 0) ? 0u : add_carry_out(A[0], x);
            MontQ q = mont_quotient(A[0]);
            redc_odd(B, q);
            q.nz += cx;
            redc_even(A, q);
            x = A[1];
#pragma unroll
            for (int k = 0; k < 7; k++) A[k] = A[k + 2];
            A[7] = t[i + 9];
            A[8] = 0;
        }
        {
            uint32_t cx = add_carry_out(B[0], x);
            MontQ q = mont_quotient(B[0]);
            redc_odd(A, q);
            q.nz += cx;
            redc_even(B, q);
            x = B[1];
#pragma unroll
            for (int k = 0; k < 7; k++) B[k] = B[k + 2];
            B[7] = t[i + 10];
            B[8] = 0;
        }
    }
    merge_even_odd(r, A, B, x);
}

// r = (y + sum_{j<N} C_j * W_j) / 2^256 mod p, identical to dot_mont_plus<N>.
//   cst(j, k)   limb k of constant j               (uniform)
//   cs(j, k)    limb k < 4 of (Cl_j + Ch_j) mod 2^128, k == 4: the carry kappa_j   (uniform, from the table)
//   word(j, i)  limb i of state word j
template <int N, class Cst, class Cs, class Word>
HADES_DEV void dot_kara_plus(uint32_t (&r)[9], Cst cst, Cs cs, Word word, const uint32_t (&y)[16]) {
    uint32_t M[9];
    {
        // WS_j = Wl_j + Wh_j (mod 2^128) with carry gamma_j; K = sum_j gamma_j*CS_j + kappa_j*WS_j (+ kappa_j*gamma_j*2^128)
        uint32_t ws[N][4], gamma[N];
        uint32_t K[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < N; j++) {
            gamma[j] = add4_carry(ws[j], word(j, 0), word(j, 1), word(j, 2), word(j, 3), word(j, 4), word(j, 5), word(j, 6), word(j, 7));
            cond_add4(K, gamma[j], cs(j, 0), cs(j, 1), cs(j, 2), cs(j, 3));
            if (cs(j, 4)) {
                add4_into5(K, ws[j][0], ws[j][1], ws[j][2], ws[j][3]);
                K[4] += gamma[j];
            }
        }
        // MM starts from Y_lo + Y_hi + K*2^128
        uint32_t ys[9];
        {
            uint32_t ylo[9], yhi[9];
#pragma unroll
            for (int k = 0; k < 8; k++) { ylo[k] = y[k]; yhi[k] = y[8 + k]; }
            ylo[8] = 0; yhi[8] = 0;
            addn_nc<9>(ys, ylo, yhi);
            uint32_t k9[9] = {0, 0, 0, 0, K[0], K[1], K[2], K[3], K[4]};
            addn_nc<9>(ys, ys, k9);
        }
        wide4_acc<N, true>(
            M, [&](int j, int k) { return cs(j, k); }, [&](int j, int i) { return ws[j][i]; }, ys);
    }
    uint32_t L[9], H[9];
    {
        uint32_t ylo[9];
#pragma unroll
        for (int k = 0; k < 8; k++) ylo[k] = y[k];
        ylo[8] = 0;
        wide4_acc<N, true>(
            L, [&](int j, int k) { return cst(j, k); }, [&](int j, int i) { return word(j, i); }, ylo);
    }
    {
        uint32_t yhi[9];
#pragma unroll
        for (int k = 0; k < 8; k++) yhi[k] = y[8 + k];
        yhi[8] = 0;
        wide4_acc<N, true>(
            H, [&](int j, int k) { return cst(j, 4 + k); }, [&](int j, int i) { return word(j, 4 + i); }, yhi);
    }
    uint32_t U[9], D[9], T[17];
    addn_nc<9>(U, L, H);
    sub9_nb(D, M, U);
    kara_assemble(T, L, D, H);
    redc17(r, T);
}

}  // namespace hades
