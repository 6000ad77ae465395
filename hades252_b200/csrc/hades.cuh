// Hades252 permutation, one state per thread, state held in registers across all 4+59+4 rounds.
//
// Mirrors the reference's algorithm skeleton (paths relative to /root/reference):
//   Strategy::perm                src/strategies.rs:140-157   4 full, 59 partial, 4 full rounds
//   apply_full_round              src/strategies.rs:107-119   ARK(all) -> x^5(all) -> MDS
//   apply_partial_round           src/strategies.rs:79-93     ARK(all) -> x^5(LAST word) -> MDS
//   ScalarStrategy::add_round_key src/strategies/scalar.rs:23-30
//   ScalarStrategy::quintic_s_box src/strategies/scalar.rs:32-34
//   ScalarStrategy::mul_matrix    src/strategies/scalar.rs:36-49
// Round r consumes ROUND_CONSTANTS[r*W .. r*W+W) (strategies.rs:141 iterator order).
//
// What differs from the reference (results are identical because F_p arithmetic is exact and the
// outputs are fully reduced): each MDS output row is ONE W-term Montgomery dot product with a single
// reduction (the reference reduces each of the W products and each of the W additions).
//
// `C` is a constants policy with static members
//     uint32_t C::ark(int idx, int limb)          limb of ROUND_CONSTANTS[idx]   (Montgomery limbs)
//     uint32_t C::mds(int row, int col, int limb) limb of MDS_MATRIX[row][col]
// which on the device read `__constant__` tables (uniform-register operands of IMAD.WIDE).
#pragma once
#include "fr.cuh"

namespace hades {

constexpr int kFullRounds = 8;      // lib.rs:22 TOTAL_FULL_ROUNDS
constexpr int kPartialRounds = 59;  // lib.rs:26 PARTIAL_ROUNDS
constexpr int kRounds = kFullRounds + kPartialRounds;

// words[j] += ROUND_CONSTANTS[base + j]
template <int W, class C>
HADES_DEV void add_round_key(Fr (&s)[W], int base) {
#pragma unroll
    for (int j = 0; j < W; j++) {
        Fr c;
#pragma unroll
        for (int k = 0; k < 8; k++) c.l[k] = C::ark(base + j, k);
        fr_add(s[j], s[j], c);
    }
}

// out[k] = sum_j MDS[k][j] * v[j]; one lazily reduced dot product per row.
// Row bound: sum < W*p^2  =>  row < p*(1 + W*p/R) = p*(1 + 0.4528 W): W<=6 -> < 4p, W<=14 -> < 8p.
template <int W, class C>
HADES_DEV void mul_matrix(Fr (&s)[W]) {
    Fr out[W];
#pragma unroll
    for (int row = 0; row < W; row++) {
        uint32_t r[9];
        dot_mont<W>(
            r, [&](int j, int k) { return C::mds(row, j, k); }, [&](int j, int i) { return s[j].l[i]; });
        canon<(W <= 6) ? 1 : 2>(out[row], r);
    }
#pragma unroll
    for (int j = 0; j < W; j++) s[j] = out[j];
}

template <int W, class C>
HADES_DEV void hades_perm(Fr (&s)[W]) {
    constexpr int kHalf = kFullRounds / 2;
#if !HADES_EMUL
#pragma unroll 1
#endif
    for (int r = 0; r < kRounds; r++) {
        add_round_key<W, C>(s, r * W);
        if (r < kHalf || r >= kHalf + kPartialRounds) {
#pragma unroll
            for (int j = 0; j < W; j++) fr_sbox(s[j]);
        } else {
            fr_sbox(s[W - 1]);
        }
        mul_matrix<W, C>(s);
    }
}


// ------------------------------------------------------------------------------------------------
// Optimised schedule: identical outputs, fewer multiplications.
//
// The partial rounds apply x^5 to the last word only (strategies.rs:83-89), so the rest of those
// rounds is linear and is re-associated on the host once (host_tables.hpp): round constants are
// pushed through the linear layer (only one scalar e_q per round remains) and each dense MDS is
// factored into a sparse matrix [[I, b], [chat^T, d]] whose dense factor migrates into the previous
// round.  Per partial round: W-1 single products  w_i += b_i*s  plus one W-term dot product, i.e.
// 2W-1 field multiplications instead of W^2.  F_p arithmetic is exact and every output is
// canonical, so the results are bit-identical to `Strategy::perm`.
//
// `T` is a table policy:  uint32_t T::tab(int entry, int limb)  over the layout of host_tables.hpp:
//   [0, 8W) full-round ARK | MDS (W*W*KR) | PRE (W*W*KR) | C4' (W) | 59 x { e, d*KD, chat[W-1]*KD, b[W-1]*KB }
// (K* = short-reduction versions per constant, OptLayout)
// ------------------------------------------------------------------------------------------------
// Short-reduction versions (fr.cuh mul_const_short / dot_mont_steps) kept per constant: the more
// versions, the fewer reduction products (6 * 8/K per reduction instead of 48) but the more constant
// memory and the larger the bound to canonicalise.  Chosen per width to fit the 64 KB constant bank.
template <int W>
struct OptLayout {
#if defined(HADES_KB) && defined(HADES_KR) && defined(HADES_KD)  // tuning overrides (tools/)
    static constexpr int kShortB = HADES_KB, kShortRow = HADES_KR, kShortDot = HADES_KD;
#else
    // Measured on B200 (W = 5): K > 1 on the multi-term rows / dot product needs more than the 63 uniform
    // registers for the constants of one step, spills, and loses 10-20 %; on the single-term b products
    // K = 4 is a clear win (+9.5 %).
    static constexpr int kShortB = (W <= 5) ? 4 : 2;  // b constants of the sparse rounds
    static constexpr int kShortRow = 1;               // dense rows (MDS, PRE) of the full rounds
    static constexpr int kShortDot = 1;               // chat / d of the sparse rounds
#endif
    static constexpr int kArk = 0;
    static constexpr int kMds = kFullRounds * W;
    static constexpr int kPre = kMds + W * W * kShortRow;
    static constexpr int kC4 = kPre + W * W * kShortRow;
    static constexpr int kSparse = kC4 + W;
    // per sparse round: e | d (kShortDot) | chat[W-1] (kShortDot each) | b[W-1] (kShortB each)
    static constexpr int kSparseD = 1;
    static constexpr int kSparseChat = kSparseD + kShortDot;
    static constexpr int kSparseB = kSparseChat + (W - 1) * kShortDot;
    static constexpr int kSparseStride = kSparseB + (W - 1) * kShortB;
    static constexpr int kEntries = kSparse + kPartialRounds * kSparseStride;
};

template <int W, class T>
HADES_DEV void add_table_vector(Fr (&s)[W], int base) {
#pragma unroll
    for (int j = 0; j < W; j++) {
        Fr c;
#pragma unroll
        for (int k = 0; k < 8; k++) c.l[k] = T::tab(base + j, k);
        fr_add(s[j], s[j], c);
    }
}

// Code-size discipline.  The kernels are bound by the integer-multiply pipe, and the profile of the
// fully unrolled version showed `no_instruction` stalls: its partial-round loop body was 42 KB of SASS,
// beyond the instruction cache.  So the per-word work is written as REAL loops over a rotating register
// file: each trip works on word 0 and then rotates the words by one (plain register moves on the
// otherwise idle ALU pipe); the loop counter only selects a (warp-uniform) constant-table offset.
#if HADES_EMUL
#define HADES_NO_UNROLL
#else
#define HADES_NO_UNROLL _Pragma("unroll 1")
#endif

#ifndef HADES_SYNC_PERIOD
#define HADES_SYNC_PERIOD 1  // block barrier every this many partial rounds (lockstep kernels; tuning knob)
#endif
struct NoSync {
    static HADES_DEV void sync() {}
};

template <int N>
HADES_DEV void rotate_in(Fr (&s)[N], const Fr& incoming) {  // s <- (s[1], ..., s[N-1], incoming)
    Fr tmp = incoming;
#pragma unroll
    for (int k = 0; k + 1 < N; k++) s[k] = s[k + 1];
    s[N - 1] = tmp;
}

// full round with a dense matrix stored at table entry `mat`
template <int W, class T>
HADES_DEV void full_round_opt(Fr (&s)[W], int ark, int mat) {
    // ARK + S-box on every word: W trips of (add constant, x^5) on word 0, rotating
    HADES_NO_UNROLL
    for (int j = 0; j < W; j++) {
        Fr c, x = s[0];
#pragma unroll
        for (int k = 0; k < 8; k++) c.l[k] = T::tab(ark + j, k);
        fr_add(x, x, c);
        fr_sbox(x);
        rotate_in<W>(s, x);
    }
    // MDS: one lazily reduced W-term dot product per row, rows pushed into a rotating output file
    Fr out[W];
#pragma unroll
    for (int j = 0; j < W; j++) out[j] = s[j];  // placeholder values, all W get overwritten
    HADES_NO_UNROLL
    for (int row = 0; row < W; row++) {
        // row = sum_j M[row][j] * s_j with the short reduction: every constant has KR versions and every
        // s_j is consumed in KR pieces: W*KR terms, 8/KR steps.  Bound: W*KR terms of (< p) * (< 2^(256/KR))
        // => < (W*KR + 1) p.
        constexpr int KR = OptLayout<W>::kShortRow;
        constexpr int kSteps = 8 / KR;
        uint32_t r[9];
        const int base = mat + row * W * KR;
        dot_mont_steps<W * KR, kSteps>(
            r, [&](int jj, int k) { return T::tab(base + jj, k); },
            [&](int jj, int i) { return s[jj / KR].l[(jj % KR) * kSteps + i]; });
        Fr res;
        canon<canon_log2_for(W * KR + 1)>(res, r);
        rotate_in<W>(out, res);
    }
#pragma unroll
    for (int j = 0; j < W; j++) s[j] = out[j];
}

// one sparse partial round; `base` = table entry of {e, d, b[W-1], chat[W-1]}
template <int W, class T, class Sync = NoSync>
HADES_DEV void partial_round_opt(Fr (&s)[W], int base) {
    constexpr int t = W - 1;
    // last word: + e_q, then x^5.  The S-box output stays lazily reduced (< 1.886p < 2^256): it only
    // feeds products below.
    {
        Fr e;
#pragma unroll
        for (int k = 0; k < 8; k++) e.l[k] = T::tab(base, k);
        fr_add(s[t], s[t], e);
    }
    Fr y, x4;
    fr_pow4_lazy(x4, s[t]);
    fr_mul_lazy(y, x4, s[t]);
#ifdef HADES_SYNC_MID
    Sync::sync();
#endif
    // new last word = sum_{j<t} chat_j * w_j + d * y  (uses the OLD w_j)
    // bound: (t + 1.886) p^2  =>  < p (1 + 0.4528 (t + 1.886)): W=5 -> 3.67p, W=9 -> 5.48p
    {
        // terms: (chat_j, w_j) for j < t and (d, y); each constant in KD short-reduction versions.
        // Bound: KD = 1: (t + 1.886) * 0.4528 + 1 (< 4p at W = 5, < 8p at W = 9); KD > 1: < (W*KD + 1) p.
        typedef OptLayout<W> L;
        constexpr int KD = L::kShortDot;
        constexpr int kSteps = 8 / KD;
        uint32_t r[9];
        dot_mont_steps<W * KD, kSteps>(
            r,
            [&](int jj, int k) {
                return (jj / KD) < t ? T::tab(base + L::kSparseChat + jj, k) : T::tab(base + L::kSparseD + (jj % KD), k);
            },
            [&](int jj, int i) {
                return (jj / KD) < t ? s[jj / KD].l[(jj % KD) * kSteps + i] : y.l[(jj % KD) * kSteps + i];
            });
        canon<(KD == 1) ? ((W <= 5) ? 1 : 2) : canon_log2_for(W * KD + 1)>(s[t], r);
    }
    // w_i += b_i * y with the SHORT reduction (fr.cuh mul_const_short): K versions of b_i, y consumed in K
    // pieces: 64 + 48/K products instead of 64 + 48.  Bound: < (K + 1) p, plus w_i < p.  t trips on word 0, rotating the first t words (back in place after t).
    Fr w[t];
#pragma unroll
    for (int i = 0; i < t; i++) w[i] = s[i];
    HADES_NO_UNROLL
    for (int i = 0; i < t; i++) {
        uint32_t q[9];
        constexpr int K = OptLayout<W>::kShortB;
        const int bi = base + OptLayout<W>::kSparseB + K * i;
        mul_const_short<K>(q, [&](int j, int k) { return T::tab(bi + j, k); }, y);
        uint32_t sum[9], q8[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; k++) q8[k] = q[k];
        uint32_t c = add8(lo, q8, w[0].l);
#pragma unroll
        for (int k = 0; k < 8; k++) sum[k] = lo[k];
        sum[8] = q[8] + c;
        Fr res;
        canon<canon_log2_for(OptLayout<W>::kShortB + 2)>(res, sum);  // < (K + 1) p + w_i
        rotate_in<t>(w, res);
    }
#pragma unroll
    for (int i = 0; i < t; i++) s[i] = w[i];
}

// `Sync::sync()` is called once per round; kernels with uniform control flow pass a block barrier so
// that the warps of a block stay in lockstep and share instruction-cache lines.
template <int W, class T, class Sync = NoSync>
HADES_DEV void hades_perm_opt(Fr (&s)[W]) {
    typedef OptLayout<W> L;
    constexpr int kHalf = kFullRounds / 2;
#if !HADES_EMUL
#pragma unroll 1
#endif
    for (int f = 0; f < kFullRounds; f++) {
        full_round_opt<W, T>(s, L::kArk + f * W, (f == kHalf - 1) ? L::kPre : L::kMds);
        Sync::sync();
        if (f == kHalf - 1) {
            add_table_vector<W, T>(s, L::kC4);
#if !HADES_EMUL
#pragma unroll 1
#endif
            for (int q = 0; q < kPartialRounds; q++) {
                partial_round_opt<W, T, Sync>(s, L::kSparse + q * L::kSparseStride);
                Sync::sync();
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Canonical-form schedule with a diagonal gauge (algo 2; derivation in host_tables.hpp, derive_tables_ccf).
// Every word is kept multiplied by a host-chosen scalar so that one matrix entry per output row is 1:
//   full rounds 0..6:   row_i = s_0 + sum_{j>=1} M^_f[i][j] s_j          (W-1 products per row)
//   partial round q:    x += e_q ; y = x^4 * x (512 bits) ; w_t' = (alpha_q . w + y) / R ; x' = (c_q . w + y) / R ;
//                       w_i' = w_{i+1}
//   P^-1 stage:         z_i = w_0 + sum_{j>=1} Q[i][j] w_j
//   full round 7:       dense (removes the gauge)
// ------------------------------------------------------------------------------------------------
template <int W>
struct CcfLayout {
    static constexpr int t = W - 1;
    static constexpr int kArk = 0;
    static constexpr int kMat = kFullRounds * W;                 // 7 matrices of W x (W-1)
    static constexpr int kMatStride = W * t;
    static constexpr int kLast = kMat + (kFullRounds - 1) * kMatStride;  // W x W
    static constexpr int kC4 = kLast + W * W;
    static constexpr int kPart = kC4 + W;                        // 59 x { e, alpha[t], c[t] }
    static constexpr int kPartStride = 2 * t + 1;
    static constexpr int kPinv = kPart + kPartialRounds * kPartStride;  // t x (t-1)
    static constexpr int kEntries = kPinv + t * (t - 1);
};

// Which operand of a constant-times-state dot product is walked by the carry chains (`vec`, 4 limbs per chain) and which
// is consumed one limb per step (`sca`)?  The product is symmetric, so this is purely a register-allocation choice.
// Constants reach IMAD.WIDE as uniform-register operands: with the constants as `vec` ALL 8 limbs of all N constants
// are wanted in uniform registers for the whole dot product (N = 8 at W = 9: 64 > the 63 that exist, so ptxas falls
// back to LDC into ordinary registers -- 379 LDC per partial round and 144 bytes of spills in the round-1 build); with
// the constants as `sca` only N uniform values are live per step and the state limbs, which sit in registers anyway,
// are the chain operands.  W <= 5 keeps the constants as `vec` (32 uniform registers, measured faster).
#ifndef HADES_SWAP_FROM_W
#define HADES_SWAP_FROM_W 9
#endif
template <int W>
struct DotRoles {
    static constexpr bool kConstIsSca = W >= HADES_SWAP_FROM_W;
};
// r = (t + sum_j tab[base + j] * word(j)) / R  (kInject) or without t; `word(j, limb)` yields state limbs
template <int N, bool kSwap, bool kInject, class T, class Word>
HADES_DEV void const_dot(uint32_t (&r)[9], int base, Word word, const uint32_t* t) {
    if constexpr (kSwap) {
        dot_mont_core<N, 8, kInject>(
            r, [&](int j, int k) { return word(j, k); }, [&](int j, int i) { return T::tab(base + j, i); }, t);
    } else {
        dot_mont_core<N, 8, kInject>(
            r, [&](int j, int k) { return T::tab(base + j, k); }, [&](int j, int i) { return word(j, i); }, t);
    }
}

// r (9 limbs) += v (8 limbs)
HADES_DEV void add_into9(uint32_t (&r)[9], const Fr& v) {
    uint32_t r8[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; k++) r8[k] = r[k];
    uint32_t c = add8(lo, r8, v.l);
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = lo[k];
    r[8] += c;
}

// out_i = s_0 + sum_{j=1..N-1} tab[base + i*(N-1) + j-1] * s_j  for i < ROWS, written into the first ROWS
// words of s (rotating output file; words ROWS..N-1 keep their values).  Inputs canonical.
// Bound: 1 + (1 + 0.4528 (N-1)):  N <= 5 -> < 4p, N <= 13 -> < 8p.
template <int N, int ROWS, class T>
HADES_DEV void unit_column_rows(Fr (&s)[N], int base) {
    static_assert(N >= 2 && N <= 13 && ROWS <= N, "bounds above");
    Fr out[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; j++) out[j] = s[j];  // placeholders, all overwritten
    HADES_NO_UNROLL
    for (int row = 0; row < ROWS; row++) {
        uint32_t r[9];
        const int b = base + row * (N - 1);
        const_dot<N - 1, DotRoles<N>::kConstIsSca, false, T>(r, b, [&](int j, int i) { return s[j + 1].l[i]; }, nullptr);
        add_into9(r, s[0]);
        Fr res;
        canon<(N <= 5) ? 1 : 2>(res, r);
        rotate_in<ROWS>(out, res);
    }
#pragma unroll
    for (int j = 0; j < ROWS; j++) s[j] = out[j];
}

// ARK + S-box on every word (W trips on word 0, rotating), constants at table entry `ark`
template <int W, class T>
HADES_DEV void ark_sbox_all(Fr (&s)[W], int ark) {
    HADES_NO_UNROLL
    for (int j = 0; j < W; j++) {
        Fr c, x = s[0];
#pragma unroll
        for (int k = 0; k < 8; k++) c.l[k] = T::tab(ark + j, k);
        fr_add(x, x, c);
        fr_sbox(x);
        rotate_in<W>(s, x);
    }
}

template <int W, class T>
HADES_DEV void partial_round_ccf(Fr (&s)[W], int base) {
    constexpr int t = W - 1;
    {
        Fr e;
#pragma unroll
        for (int k = 0; k < 8; k++) e.l[k] = T::tab(base, k);
        fr_add(s[t], s[t], e);
    }
    // y = x^4 * x is NOT reduced on its own: the 512-bit product enters both dot products below as an addend
    // and shares their reductions (the gauge made its coefficient 1 in both).
    // Bound: (sum_{j<t} c_j w_j + x^4 x) / R + p  <  p (1 + 0.4528 (t + 1.956)):  W=3: 2.8p, W=5: 3.7p, W=9: 5.6p
    Fr x4;
    fr_pow4_lazy(x4, s[t]);
    uint32_t y[16];
    mul_wide(y, x4.l, s[t].l);
    Fr newx, neww;
    {
        uint32_t r[9];
        const_dot<t, DotRoles<W>::kConstIsSca, true, T>(r, base + 1 + t, [&](int j, int i) { return s[j].l[i]; }, y);
        canon<(W <= 5) ? 1 : 2>(newx, r);
    }
    {
        uint32_t r[9];
        const_dot<t, DotRoles<W>::kConstIsSca, true, T>(r, base + 1, [&](int j, int i) { return s[j].l[i]; }, y);
        canon<(W <= 5) ? 1 : 2>(neww, r);
    }
    // shift the words, the new one enters at the end
#pragma unroll
    for (int i = 0; i + 1 < t; i++) s[i] = s[i + 1];
    s[t - 1] = neww;
    s[t] = newx;
}

template <int W, class T, class Sync = NoSync>
HADES_DEV void hades_perm_ccf(Fr (&s)[W]) {
    typedef CcfLayout<W> L;
    constexpr int kHalf = kFullRounds / 2;
    constexpr int t = W - 1;
#if !HADES_EMUL
#pragma unroll 1
#endif
    for (int f = 0; f + 1 < kFullRounds; f++) {
        ark_sbox_all<W, T>(s, L::kArk + f * W);
        unit_column_rows<W, W, T>(s, L::kMat + f * L::kMatStride);
        Sync::sync();
        if (f == kHalf - 1) {
            add_table_vector<W, T>(s, L::kC4);
#if !HADES_EMUL
#pragma unroll 1
#endif
            for (int q = 0; q < kPartialRounds; q++) {
                partial_round_ccf<W, T>(s, L::kPart + q * L::kPartStride);
#if HADES_SYNC_PERIOD > 1
                if (q % HADES_SYNC_PERIOD == HADES_SYNC_PERIOD - 1) Sync::sync();
#else
                Sync::sync();
#endif
            }
            // back to the original basis (gauged): z_i = w_0 + sum_{j>=1} Q[i][j] w_j ; the last word stays
            Fr w[t];
#pragma unroll
            for (int i = 0; i < t; i++) w[i] = s[i];
            unit_column_rows<t, t, T>(w, L::kPinv);
#pragma unroll
            for (int i = 0; i < t; i++) s[i] = w[i];
        }
    }
    // last full round: dense rows, lands on the true state.  Rows 0..W-2 go through a rotating file of
    // W-1 words and the last row is peeled, so at most 2W-1 words are live (register budget: 96 at W = 5).
    ark_sbox_all<W, T>(s, L::kArk + (kFullRounds - 1) * W);
    {
        auto dense_row = [&](Fr& res, int b) {
            uint32_t r[9];
            const_dot<W, DotRoles<W>::kConstIsSca, false, T>(r, b, [&](int j, int i) { return s[j].l[i]; }, nullptr);
            canon<(W <= 6) ? 1 : 2>(res, r);
        };
        Fr out[t];
#pragma unroll
        for (int j = 0; j < t; j++) out[j] = s[j];
        HADES_NO_UNROLL
        for (int row = 0; row < t; row++) {
            Fr res;
            dense_row(res, L::kLast + row * W);
            rotate_in<t>(out, res);
        }
        Fr last;
        dense_row(last, L::kLast + t * W);
#pragma unroll
        for (int j = 0; j < t; j++) s[j] = out[j];
        s[t] = last;
    }
    Sync::sync();
}

}  // namespace hades
