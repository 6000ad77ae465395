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

}  // namespace hades
