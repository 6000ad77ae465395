// Cooperative (low-latency) Hades252 permutation: ONE width-5 state per group of 8 lanes, or per warp.
//
// Why: the one-thread-per-state kernels are throughput kernels.  A lone warp needs ~270 us for its 74 200
// dependent-ish limb products, so every batch below ~2^14 states -- a lone `Strategy::perm`
// (src/strategies.rs:140), the top levels of a Merkle tree -- pays that floor.  Lanes of a warp are free when
// the batch is small, so here a group of G lanes shares one state (BASELINE north_star: "one thread (or a
// small cooperative group) per state"): G = 8 (kCoopLanes) up to a few thousand states, G = 32 (kCoopWide) for the
// smallest batches.
//
// How: SIMT lanes must run the same instruction, so the work is cut into SLOTS of one Montgomery product
// r = a*b/R per lane (fr.cuh dot_mont<1>, CIOS) with lane-dependent operands, glued by warp shuffles:
//   * the state is REPLICATED in all lanes of the group (canonical words), so operands are picked by lane
//     index and nothing has to be routed before a slot;
//   * partial round (gauged canonical form, hades.cuh partial_round_ccf): the S-box chain x^2, x^4, x^5 runs
//     on lane 0 in slots 0..2; the 8 products of the two dot products (c_q.w and alpha_q.w) ride along in the
//     other lanes (G = 8: 7 in slot 0, 1 in slot 1; G = 32: all in slot 0, which makes slot 1 a pure squaring
//     slot); their sums are formed by xor-butterflies and canonicalised while lane 0 is still multiplying, so only
//     "+ x^5, two conditional subtractions" is left after slot 2: 3 products deep instead of 840/112 = 7.5;
//   * full round: x^5 of the 5 words on lanes 0..4 (two squaring slots, one product slot), then the 20 (25 in the
//     last, dense round) matrix products in 3 (4 + 1) slots with 8 lanes, in 1 (2) with a warp; each row is summed by
//     a butterfly over 4 lanes;
//   * every product is reduced on its own (Montgomery reduction is linear), sums are taken on the reduced
//     9-limb values and canonicalised once per output word.
// Same tables as the one-thread canonical-form kernel (CcfLayout<5>), same F_p values, canonical outputs:
// bit-identical results.  Slots per permutation: 59*3 + 7*6 + 7 + 2 = 228 with 8 lanes, 59*3 + 7*4 + 5 + 1 = 211
// with a warp (59 + 16 of them squarings).
//
// `T::tab(entry, limb)` must accept a lane-dependent entry (the kernels copy the table to shared memory).
// Shuffles go through coop_shfl_n<G> / coop_shfl_xor_n<G>: warp shuffles of width G on the device; in the host
// emulation build (tests/host_emul) G threads exchange through a barrier.
#pragma once
#include "hades.cuh"

namespace hades {

constexpr int kCoopLanes = 8;

#if HADES_EMUL
// provided by the emulation harness: lane id of the calling thread and the exchange primitives
int coop_emul_lane();
void coop_emul_exchange(uint32_t* v, int n, int src_lane_xor, int src_lane_abs);  // abs < 0: use xor
template <int G = kCoopLanes>
inline void coop_shfl_xor_n(uint32_t* v, int n, int mask) { coop_emul_exchange(v, n, mask, -1); }
template <int G = kCoopLanes>
inline void coop_shfl_n(uint32_t* v, int n, int src) { coop_emul_exchange(v, n, 0, src); }
#else
template <int G = kCoopLanes>
__device__ __forceinline__ void coop_shfl_xor_n(uint32_t* v, int n, int mask) {
#pragma unroll
    for (int k = 0; k < n; k++) v[k] = __shfl_xor_sync(0xffffffffu, v[k], mask, G);
}
template <int G = kCoopLanes>
__device__ __forceinline__ void coop_shfl_n(uint32_t* v, int n, int src) {
#pragma unroll
    for (int k = 0; k < n; k++) v[k] = __shfl_sync(0xffffffffu, v[k], src, G);
}
#endif
// The same permutation with a whole WARP per state (G = kCoopWide): the 20 matrix products of a full round fit one
// slot (3 with 8 lanes), all 8 dot products of a partial round fit its first slot, which frees the second slot to
// be a pure squaring (84 products instead of 112).  A quarter of the capacity per wave (592 states on 148 SMs), so it
// only serves the smallest batches: a lone `Strategy::perm`, the top levels of a Merkle tree.
constexpr int kCoopWide = 32;

// dst = cond ? a : dst  (per limb; SEL on the device)
template <int N>
HADES_DEV void coop_pick(uint32_t (&dst)[N], bool cond, const uint32_t (&a)[N]) {
#pragma unroll
    for (int k = 0; k < N; k++) dst[k] = cond ? a[k] : dst[k];
}

// one slot: r = a*b/R mod p, 8 limbs, r < p + a*b/R (every product in this file has a < 1.96p and b < 1.96p, so
// r < 2^256 and limb 8 is zero).  One CIOS stream (dot_mont<1>, 112 products).  Measured and rejected: splitting the
// product into two independent carry-chain streams (low and high limbs of b reduced separately, 136 products) to
// double the instruction-level parallelism -- a lone warp is NOT latency-bound but at ~57 % of the FMA pipe already
// (ncu), so the extra products cost more than the parallelism buys: 166 us instead of 125 us per permutation
// (profiles/r02_latency_coop_split_ab.txt).
HADES_DEV void coop_mmul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t t[9];
    dot_mont<1>(t, [&](int, int k) { return a[k]; }, [&](int, int i) { return b[i]; });
    HADES_ASSERT(t[8] == 0);
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = t[k];
}

// v (9 limbs) += a (8 limbs)
HADES_DEV void coop_acc8(uint32_t (&v)[9], const uint32_t (&a)[8]) {
    uint32_t lo[8], v8[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v8[k] = v[k];
    const uint32_t c = add8(lo, v8, a);
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = lo[k];
    v[8] += c;
}
// v (9 limbs) += a (9 limbs)
HADES_DEV void coop_acc9(uint32_t (&v)[9], const uint32_t (&a)[9]) {
    uint32_t a8[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a8[k] = a[k];
    coop_acc8(v, a8);
    v[8] += a[8];
}
// v += v of lane (lane ^ mask)
template <int G = kCoopLanes>
HADES_DEV void coop_butterfly(uint32_t (&v)[9], int mask) {
    uint32_t o[9];
#pragma unroll
    for (int k = 0; k < 9; k++) o[k] = v[k];
    coop_shfl_xor_n<G>(o, 9, mask);
    coop_acc9(v, o);
}

template <class T>
HADES_DEV void coop_load_tab(uint32_t (&c)[8], int entry) {
#if HADES_EMUL
    for (int k = 0; k < 8; k++) c[k] = T::tab(entry, k);
#else
    const uint4* p = T::ptr4(entry);  // lane-dependent entry: two 128-bit shared-memory loads
    const uint4 lo = p[0], hi = p[1];
    c[0] = lo.x; c[1] = lo.y; c[2] = lo.z; c[3] = lo.w;
    c[4] = hi.x; c[5] = hi.y; c[6] = hi.z; c[7] = hi.w;
#endif
}

// word idx (lane-dependent, 0..N-1) of a replicated state
template <int N>
HADES_DEV void coop_select_word(uint32_t (&out)[8], const Fr (&s)[N], int idx) {
#pragma unroll
    for (int k = 0; k < 8; k++) out[k] = s[0].l[k];
#pragma unroll
    for (int j = 1; j < N; j++) coop_pick(out, idx == j, s[j].l);
}

// one squaring slot: r = a*a/R mod p, r < p + a*a/R.  Every lane squares, so the dedicated squaring (84 products
// instead of 112) can be used -- only where ALL lanes of the slot square (the S-boxes of the full rounds).
HADES_DEV void coop_msqr(uint32_t (&r)[8], const uint32_t (&a)[8]) {
    uint32_t t[9];
    sqr_mont(t, a);
    HADES_ASSERT(t[8] == 0);
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = t[k];
}

// x^5 of a canonical x, canonical result (three slots)
HADES_DEV void coop_sbox(uint32_t (&x5)[8], const uint32_t (&x)[8]) {
    uint32_t x2[8], x4[8];
    coop_msqr(x2, x);        // < 1.453 p
    coop_msqr(x4, x2);       // < 1.956 p
    coop_mmul(x5, x4, x);    // < 1.886 p
    cond_sub_p8(x5);
}

// Rows of a matrix over a group of 4 lanes: lane (g, c) with g = lane >> 2, c = lane & 3 multiplies its operand
// `b` (the same for both groups) by entry `e0 + row(g) * stride + c` and the four products plus `col0` (added by
// c == 0) are summed by a butterfly.  Returns the canonical row in every lane of the group.
// Bound: 4 products < 1.4528 p each (canonical operands) + col0 < 1.4528 p: < 7.3 p < 8 p.
template <class T, int G = kCoopLanes>
HADES_DEV void coop_row_slot(Fr& out, int lane, int entry, bool active, const uint32_t (&b)[8], const uint32_t (&col0)[8]) {
    uint32_t a[8], r[8], v[9];
    coop_load_tab<T>(a, entry);
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = active ? a[k] : 0u;
    coop_mmul(r, a, b);
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = r[k];
    v[8] = 0;
    uint32_t z[8];
#pragma unroll
    for (int k = 0; k < 8; k++) z[k] = ((lane & 3) == 0) ? col0[k] : 0u;
    coop_acc8(v, z);
    coop_butterfly<G>(v, 1);
    coop_butterfly<G>(v, 2);
    canon<2>(out, v);
}

// ARK + x^5 on all five words (lane j < 5 owns word j), then out = M * sbox: unit-column rows (full rounds 0..6:
// out_i = s_0 + sum_{j>=1} tab[mat + 4 i + j - 1] s_j) or dense rows (last round: out_i = sum_j tab[mat + 5 i + j] s_j).
template <class T, bool kDense, int G = kCoopLanes>
HADES_DEV void coop_full_round(Fr (&s)[5], int lane, int ark, int mat) {
    const int widx = lane < 5 ? lane : 0;
    uint32_t x[8], c[8], x5[8];
    coop_select_word<5>(x, s, widx);
    coop_load_tab<T>(c, ark + widx);
    {
        Fr xf, cf;
#pragma unroll
        for (int k = 0; k < 8; k++) { xf.l[k] = x[k]; cf.l[k] = c[k]; }
        fr_add(xf, xf, cf);
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = xf.l[k];
    }
    coop_sbox(x5, x);
    // operand of the row slots: word (lane & 3) + 1; column-0 term: s_0 itself (unit column) or M[row][0] * s_0
    uint32_t b[8], s0[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { b[k] = x5[k]; s0[k] = x5[k]; }
    coop_shfl_n<G>(b, 8, (lane & 3) + 1);
    coop_shfl_n<G>(s0, 8, 0);
    const int c4 = lane & 3, g = lane >> 2;
    uint32_t d[8];  // dense: lane i < 5 holds M[i][0] * s_0
    if constexpr (kDense) {
        uint32_t a[8];
        coop_load_tab<T>(a, mat + 5 * widx);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = lane < 5 ? a[k] : 0u;
        coop_mmul(d, a, s0);
    }
    constexpr int kStride = kDense ? 5 : 4, kFirst = kDense ? 1 : 0;
    auto row = [&](Fr& out, int r, bool active) {
        uint32_t col0[8];
#pragma unroll
        for (int k = 0; k < 8; k++) col0[k] = kDense ? d[k] : s0[k];
        if constexpr (kDense) coop_shfl_n<G>(col0, 8, r);
        coop_row_slot<T, G>(out, lane, mat + kStride * r + kFirst + c4, active, b, col0);
    };
    Fr t;
    if constexpr (G == kCoopWide) {
        // all five rows in ONE slot: row g on lanes 4g .. 4g+3 (g < 5), the other lanes idle
        Fr out;
        row(out, g < 5 ? g : 4, g < 5);
#pragma unroll
        for (int i = 0; i < 5; i++) { t = out; coop_shfl_n<G>(t.l, 8, 4 * i); s[i] = t; }
    } else {
        Fr outA, outB, outC;
        row(outA, g, true);      // rows 0 | 1
        row(outB, 2 + g, true);  // rows 2 | 3
        row(outC, 4, true);      // row 4 in both groups
        t = outA; coop_shfl_n<G>(t.l, 8, 0); s[0] = t;
        t = outA; coop_shfl_n<G>(t.l, 8, 4); s[1] = t;
        t = outB; coop_shfl_n<G>(t.l, 8, 0); s[2] = t;
        t = outB; coop_shfl_n<G>(t.l, 8, 4); s[3] = t;
        s[4] = outC;
    }
}

// partial round q (table entries base = {e, alpha[4], c[4]}): see the file header.
// The round constant is folded one round ahead: on entry s[4] already contains e_q (the caller adds e_0), and the
// round adds e_{q+1} (table entry `next_e`, < 0 after the last round) to the new x.  Everything that does not depend on
// the S-box output is finished while lane 0 is still multiplying: the two dot-product sums (c.w + e_{q+1} in lanes 0..3,
// alpha.w in lanes 4..7) are canonicalised before slot 2, so that after slot 2 only  "+ x^5, two conditional
// subtractions" is left on the critical path, once per lane half; the halves then swap their results.
template <class T>
HADES_DEV void coop_partial_round(Fr (&s)[5], int lane, int base, int next_e) {
    // slot 0: lane 0: x*x | lanes 1..3: c_j * w_j (j = lane - 1) | lanes 4..7: alpha_j * w_j (j = lane - 4)
    const int j0 = lane == 0 ? 4 : (lane < 4 ? lane - 1 : lane - 4);
    uint32_t a[8], b[8], r0[8], r1[8], r2[8];
    coop_select_word<5>(b, s, j0);
    coop_load_tab<T>(a, base + (lane < 4 ? 5 + j0 : 1 + j0));  // lane 0 reads a valid but unused entry
    coop_pick(a, lane == 0, s[4].l);
    coop_mmul(r0, a, b);
    // slot 1: lane 0: x^2 * x^2 | lane 1: c_3 * w_3 | others idle
    coop_load_tab<T>(a, base + 5 + 3);
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = lane == 1 ? a[k] : 0u; b[k] = s[3].l[k]; }
    coop_pick(a, lane == 0, r0);
    coop_pick(b, lane == 0, r0);
    coop_mmul(r1, a, b);
    // slot 2: lane 0: x^4 * x (others idle).  Written BEFORE the sums below, which do not depend on it: the sums,
    // butterflies and the first canonicalisation are meant to fill the issue slots between its multiplies.
    {
        uint32_t a2[8], b2[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { a2[k] = lane == 0 ? r1[k] : 0u; b2[k] = s[4].l[k]; }
        coop_mmul(r2, a2, b2);
    }
    // sums of the reduced products (off the S-box chain): lanes 0..3 -> c . w, lanes 4..7 -> alpha . w
    uint32_t v[9];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = lane == 0 ? 0u : r0[k];
    v[8] = 0;
    {
        uint32_t z[8];
#pragma unroll
        for (int k = 0; k < 8; k++) z[k] = lane == 1 ? r1[k] : 0u;
        coop_acc8(v, z);
    }
    coop_butterfly(v, 1);
    coop_butterfly(v, 2);
    // lanes 0..3: + e_{q+1}; then canonical (< 4 * 1.4528 p + p < 8 p), still off the critical path
    if (next_e >= 0) {
        uint32_t e[8];
#pragma unroll
        for (int k = 0; k < 8; k++) e[k] = lane < 4 ? T::tab(next_e, k) : 0u;
        coop_acc8(v, e);
    }
    Fr pre;
    canon<2>(pre, v);
    coop_shfl_n(r2, 8, 0);  // y = x^5 / gauge, < 1.886 p, to every lane
    uint32_t sum[9];
#pragma unroll
    for (int k = 0; k < 8; k++) sum[k] = pre.l[k];
    sum[8] = 0;
    coop_acc8(sum, r2);  // < 2.886 p
    Fr mine, other;
    canon<1>(mine, sum);  // lanes 0..3: the new x (+ e_{q+1}), lanes 4..7: the new w
    other = mine;
    coop_shfl_xor_n(other.l, 8, 4);
#pragma unroll
    for (int i = 0; i < 3; i++) s[i] = s[i + 1];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        s[3].l[k] = lane < 4 ? other.l[k] : mine.l[k];
        s[4].l[k] = lane < 4 ? mine.l[k] : other.l[k];
    }
}

// the same round with a warp per state: slot 0: lane 0: x*x | lanes 4..7: c_j * w_j | lanes 8..11: alpha_j * w_j;
// slot 1: a pure SQUARING slot (lane 0: x^2 -> x^4, the other lanes square their own product and drop it);
// slot 2: lane 0: x^4 * x.  Sums, folded round constant and canonicalisation as in the 8-lane round.
template <class T>
HADES_DEV void coop_partial_round_wide(Fr (&s)[5], int lane, int base, int next_e) {
    constexpr int G = kCoopWide;
    const int j0 = lane == 0 ? 4 : (lane & 3);
    const bool dot = lane >= 4 && lane < 12;
    uint32_t a[8], b[8], r0[8], r1[8], r2[8];
    coop_select_word<5>(b, s, j0);
    coop_load_tab<T>(a, base + ((lane & 8) ? 1 : 5) + (lane & 3));  // lanes outside 4..11 read a valid but unused entry
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = dot ? a[k] : 0u;
    coop_pick(a, lane == 0, s[4].l);
    coop_mmul(r0, a, b);   // every lane: < 1.4528 p (canonical operands) or 0
    coop_msqr(r1, r0);     // lane 0: x^4 < 1.956 p
    {
        uint32_t a2[8], b2[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { a2[k] = lane == 0 ? r1[k] : 0u; b2[k] = s[4].l[k]; }
        coop_mmul(r2, a2, b2);
    }
    uint32_t v[9];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = dot ? r0[k] : 0u;
    v[8] = 0;
    coop_butterfly<G>(v, 1);
    coop_butterfly<G>(v, 2);  // lanes 4..7: c . w, lanes 8..11: alpha . w (< 4 * 1.4528 p)
    if (next_e >= 0) {
        uint32_t e[8];
#pragma unroll
        for (int k = 0; k < 8; k++) e[k] = (lane & 12) == 4 ? T::tab(next_e, k) : 0u;
        coop_acc8(v, e);
    }
    Fr pre;
    canon<2>(pre, v);
    coop_shfl_n<G>(r2, 8, 0);
    uint32_t sum[9];
#pragma unroll
    for (int k = 0; k < 8; k++) sum[k] = pre.l[k];
    sum[8] = 0;
    coop_acc8(sum, r2);  // < 2.886 p
    Fr mine, nx, nw;
    canon<1>(mine, sum);  // lanes 4..7: the new x (+ e_{q+1}), lanes 8..11: the new w
    nx = mine; coop_shfl_n<G>(nx.l, 8, 4);
    nw = mine; coop_shfl_n<G>(nw.l, 8, 8);
#pragma unroll
    for (int i = 0; i < 3; i++) s[i] = s[i + 1];
    s[3] = nw;
    s[4] = nx;
}

// back to the original basis after the partial rounds: z_i = w_0 + sum_{j=1..3} tab[pinv + 3 i + j - 1] w_j (i < 4)
template <class T, int G = kCoopLanes>
HADES_DEV void coop_pinv_stage(Fr (&s)[5], int lane, int pinv) {
    const int c4 = lane & 3, g = lane >> 2;
    uint32_t b[8], w0[8];
    coop_select_word<5>(b, s, c4);  // lane c4 = 0 multiplies nothing: its entry index below is clamped and masked
#pragma unroll
    for (int k = 0; k < 8; k++) w0[k] = s[0].l[k];
    Fr t;
    if constexpr (G == kCoopWide) {
        Fr out;  // all four rows in one slot: row g on lanes 4g .. 4g+3 (g < 4)
        coop_row_slot<T, G>(out, lane, pinv + 3 * (g < 4 ? g : 3) + (c4 ? c4 - 1 : 0), c4 != 0 && g < 4, b, w0);
#pragma unroll
        for (int i = 0; i < 4; i++) { t = out; coop_shfl_n<G>(t.l, 8, 4 * i); s[i] = t; }
    } else {
        Fr outA, outB;
        coop_row_slot<T, G>(outA, lane, pinv + 3 * g + (c4 ? c4 - 1 : 0), c4 != 0, b, w0);
        coop_row_slot<T, G>(outB, lane, pinv + 3 * (2 + g) + (c4 ? c4 - 1 : 0), c4 != 0, b, w0);
        t = outA; coop_shfl_n<G>(t.l, 8, 0); s[0] = t;
        t = outA; coop_shfl_n<G>(t.l, 8, 4); s[1] = t;
        t = outB; coop_shfl_n<G>(t.l, 8, 0); s[2] = t;
        t = outB; coop_shfl_n<G>(t.l, 8, 4); s[3] = t;
    }
}

// `Strategy::perm` (src/strategies.rs:140-157) on the replicated state of one group of G lanes (8, or a whole warp)
template <class T, int G = kCoopLanes>
HADES_DEV void hades_perm_coop(Fr (&s)[5], int lane) {
    static_assert(G == kCoopLanes || G == kCoopWide, "8 lanes or a warp per state");
    typedef CcfLayout<5> L;
    constexpr int kHalf = kFullRounds / 2;
#if !HADES_EMUL
#pragma unroll 1
#endif
    for (int f = 0; f + 1 < kFullRounds; f++) {
        coop_full_round<T, false, G>(s, lane, L::kArk + f * 5, L::kMat + f * L::kMatStride);
        if (f == kHalf - 1) {
            add_table_vector<5, T>(s, L::kC4);
            {
                Fr e;
#pragma unroll
                for (int k = 0; k < 8; k++) e.l[k] = T::tab(L::kPart, k);
                fr_add(s[4], s[4], e);
            }
#if !HADES_EMUL
#pragma unroll 1
#endif
            for (int q = 0; q < kPartialRounds; q++) {
                const int base = L::kPart + q * L::kPartStride;
                const int next_e = q + 1 < kPartialRounds ? base + L::kPartStride : -1;
                if constexpr (G == kCoopWide) coop_partial_round_wide<T>(s, lane, base, next_e);
                else coop_partial_round<T>(s, lane, base, next_e);
            }
            coop_pinv_stage<T, G>(s, lane, L::kPinv);
        }
    }
    coop_full_round<T, true, G>(s, lane, L::kArk + (kFullRounds - 1) * 5, L::kLast);
}

}  // namespace hades
