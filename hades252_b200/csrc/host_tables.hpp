// Host-side derivation (once, inside hades_init) of the constant tables the kernels consume, from the
// two tables the reference owns: ROUND_CONSTANTS (src/round_constants.rs:29-48) and MDS_MATRIX
// (src/mds_matrix.rs:18-40).  This is constant preprocessing, not a data path: no state is ever
// permuted on the host.
//
// The 59 partial rounds of `Strategy::perm` (src/strategies.rs:149-151) apply the S-box to the LAST
// word only (strategies.rs:83-89), so everything else in those rounds is linear and can be
// re-associated without changing any output (F_p arithmetic is exact):
//   (1) round-constant pushing: write the state entering partial round q as z_q + d_q with
//       d_q[last] = 0; the S-box ignores d_q, so d_{q+1} = first W-1 entries of (M d_q + c_{q+1}) and
//       only the scalar e_{q+1} = (M d_q + c_{q+1})[last] has to be added to the live state.  What is
//       left after the last partial round (M d_58) merges into the next full round's constants.
//   (2) sparse factorisation: a dense D = [[A, b], [c^T, d]] (A: (W-1)x(W-1)) factors as
//       D = S * M'  with  M' = blockdiag(A, 1)  and  S = [[I, b], [c^T A^-1, d]].  M' commutes with the
//       partial S-box (it fixes the last word), so it migrates into the previous round's matrix:
//       D_58 = M, D_{q-1} = M'_q * M.  The leftover M'_0 merges into the last of the first four full
//       rounds (matrix PRE = M'_0 * M, and the first partial ARK becomes M'_0 * c_4).
// Each partial round then costs 2W-1 field multiplications instead of W^2.
//
// Table layout (entries of 4 u64 Montgomery limbs), W = width, F = 8 full rounds, Q = 59:
//   [0,            F*W)        ARK of the full rounds: rounds 0..3, then c63' = c_63 + M d_58, c_64..c_66
//   [F*W,          +W*W*KR)    MDS   (dense M, row-major, KR short-reduction versions per entry)
//   [..,           +W*W*KR)    PRE   (dense M'_0 * M, used by full round 3; same)
//   [..,           +W)         C4'   (M'_0 * c_4, added to every word before the first partial round)
//   [..,           +Q*stride)  per partial round q: e_q | d_q (KD versions) | chat_q[0..W-2] (KD each) |
//                              b_q[0..W-2] (KB each).  "K versions" of a constant c are the short-reduction
//                              constants X_j = c * 2^(256 (j+1)/K - 256), j = 0..K-1 (fr.cuh dot_mont_steps):
//                              the product is then reduced in 8/K steps instead of 8.  K per width is chosen
//                              to fit the 64 KB constant bank (short_b / short_row / short_dot below).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace hades_host {

typedef unsigned __int128 u128;
struct F {
    uint64_t l[4];
};

static const uint64_t kMod[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                 0x73eda753299d7d48ULL};
static const uint64_t kInv64 = 0xfffffffeffffffffULL;  // -p^-1 mod 2^64
static const F kOne = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL,
                        0x1824b159acc5056fULL}};  // R mod p
static const F kZero = {{0, 0, 0, 0}};

inline bool geq_mod(const uint64_t a[4], uint64_t top) {
    if (top) return true;
    for (int i = 3; i >= 0; i--) {
        if (a[i] > kMod[i]) return true;
        if (a[i] < kMod[i]) return false;
    }
    return true;
}
inline void sub_mod_inplace(uint64_t a[4]) {
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - kMod[i] - bw;
        a[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
    }
}
inline F add(const F& a, const F& b) {
    F r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_mod(r.l, (uint64_t)c)) sub_mod_inplace(r.l);
    return r;
}
inline F neg(const F& a) {
    bool zero = !(a.l[0] | a.l[1] | a.l[2] | a.l[3]);
    if (zero) return a;
    F r;
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)kMod[i] - a.l[i] - bw;
        r.l[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
    }
    return r;
}
inline F sub(const F& a, const F& b) { return add(a, neg(b)); }
inline F mul(const F& a, const F& b) {  // Montgomery product a*b/R
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * kInv64;
        c = (u128)m * kMod[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * kMod[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    F r = {{t[0], t[1], t[2], t[3]}};
    if (geq_mod(r.l, t[4])) sub_mod_inplace(r.l);
    return r;
}
inline bool is_zero(const F& a) { return !(a.l[0] | a.l[1] | a.l[2] | a.l[3]); }
inline F inv(const F& a) {  // a^(p-2), Montgomery domain in and out
    F r = kOne, base = a;
    uint64_t e[4] = {kMod[0] - 2, kMod[1], kMod[2], kMod[3]};
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, base);
        base = mul(base, base);
    }
    return r;
}

typedef std::vector<F> Vec;
typedef std::vector<Vec> Mat;

inline Vec matvec(const Mat& A, const Vec& v) {
    Vec r(A.size(), kZero);
    for (size_t i = 0; i < A.size(); i++)
        for (size_t j = 0; j < v.size(); j++) r[i] = add(r[i], mul(A[i][j], v[j]));
    return r;
}
inline Mat matmul(const Mat& A, const Mat& B) {
    Mat r(A.size(), Vec(B[0].size(), kZero));
    for (size_t i = 0; i < A.size(); i++)
        for (size_t j = 0; j < B[0].size(); j++)
            for (size_t k = 0; k < B.size(); k++) r[i][j] = add(r[i][j], mul(A[i][k], B[k][j]));
    return r;
}
// Gauss-Jordan; returns false if singular.
inline bool invert(const Mat& A, Mat& out) {
    size_t n = A.size();
    Mat M(n, Vec(2 * n, kZero));
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < n; j++) M[i][j] = A[i][j];
        M[i][n + i] = kOne;
    }
    for (size_t c = 0; c < n; c++) {
        size_t piv = c;
        while (piv < n && is_zero(M[piv][c])) piv++;
        if (piv == n) return false;
        std::swap(M[c], M[piv]);
        F iv = inv(M[c][c]);
        for (auto& x : M[c]) x = mul(x, iv);
        for (size_t r = 0; r < n; r++) {
            if (r == c || is_zero(M[r][c])) continue;
            F f = M[r][c];
            for (size_t k = 0; k < 2 * n; k++) M[r][k] = sub(M[r][k], mul(f, M[c][k]));
        }
    }
    out.assign(n, Vec(n));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) out[i][j] = M[i][n + j];
    return true;
}

constexpr int kFull = 8, kPartial = 59, kHalf = 4;

// short-reduction versions per constant; must match OptLayout<W> in hades.cuh (checked in the emulation test)
inline int short_b(int W) { return W <= 5 ? 4 : 2; }
inline int short_row(int) { return 1; }
inline int short_dot(int) { return 1; }
inline size_t sparse_stride(int W) { return 1 + (size_t)short_dot(W) * W + (size_t)short_b(W) * (W - 1); }
inline size_t table_entries(int W) {
    return (size_t)kFull * W + 2 * (size_t)W * W * short_row(W) + W + (size_t)kPartial * sparse_stride(W);
}

// ark: >= 67*W entries, mds: W*W entries (Montgomery limbs).  out: table_entries(W)*4 u64.
inline bool derive_tables(int W, const uint64_t* ark, const uint64_t* mds, std::vector<uint64_t>& out) {
    const int t = W - 1;
    auto getF = [](const uint64_t* p) { F f; for (int i = 0; i < 4; i++) f.l[i] = p[i]; return f; };
    Mat M(W, Vec(W));
    for (int i = 0; i < W; i++)
        for (int j = 0; j < W; j++) M[i][j] = getF(mds + (size_t)(i * W + j) * 4);
    std::vector<Vec> c(kFull + kPartial, Vec(W));
    for (int r = 0; r < kFull + kPartial; r++)
        for (int j = 0; j < W; j++) c[r][j] = getF(ark + (size_t)(r * W + j) * 4);
    // (1) constant pushing, forward
    Vec d(W, kZero), e(kPartial, kZero);
    for (int q = 0; q + 1 < kPartial; q++) {
        Vec u = matvec(M, d);
        for (int j = 0; j < W; j++) u[j] = add(u[j], c[kHalf + q + 1][j]);
        e[q + 1] = u[t];
        d = u;
        d[t] = kZero;
    }
    Vec tail = matvec(M, d);
    // (2) sparse factorisation, backward
    struct Sparse { Vec b, chat; F dd; };
    std::vector<Sparse> sp(kPartial);
    Mat D = M, Mp;
    for (int q = kPartial - 1; q >= 0; q--) {
        Mat A(t, Vec(t)), Ai;
        for (int i = 0; i < t; i++)
            for (int j = 0; j < t; j++) A[i][j] = D[i][j];
        if (!invert(A, Ai)) return false;
        sp[q].b.resize(t);
        sp[q].chat.assign(t, kZero);
        for (int i = 0; i < t; i++) sp[q].b[i] = D[i][t];
        for (int j = 0; j < t; j++)
            for (int k = 0; k < t; k++) sp[q].chat[j] = add(sp[q].chat[j], mul(D[t][k], Ai[k][j]));
        sp[q].dd = D[t][t];
        Mp.assign(W, Vec(W, kZero));
        for (int i = 0; i < t; i++)
            for (int j = 0; j < t; j++) Mp[i][j] = A[i][j];
        Mp[t][t] = kOne;
        D = matmul(Mp, M);
    }
    Vec c4 = matvec(Mp, c[kHalf]);
    out.clear();
    out.reserve(table_entries(W) * 4);
    auto put = [&](const F& f) { for (int i = 0; i < 4; i++) out.push_back(f.l[i]); };
    for (int r = 0; r < kHalf; r++)
        for (int j = 0; j < W; j++) put(c[r][j]);
    for (int j = 0; j < W; j++) put(add(c[kHalf + kPartial][j], tail[j]));
    for (int r = kHalf + kPartial + 1; r < kFull + kPartial; r++)
        for (int j = 0; j < W; j++) put(c[r][j]);
    // K short-reduction versions of a constant c: X_j = c * 2^(256 (j+1)/K - 256), j = 0..K-1, i.e. X_{K-1} = c
    // and each step down multiplies by 2^(-256/K) = mul(., 2^(256 - 256/K) as a plain integer).
    auto put_versions = [&](const F& cst, int K) {
        F down = kZero;
        if (K == 4) down.l[3] = 1;       // 2^192
        else if (K == 2) down.l[2] = 1;  // 2^128
        F ver[4];
        ver[K - 1] = cst;
        for (int j = K - 2; j >= 0; j--) ver[j] = mul(ver[j + 1], down);
        for (int j = 0; j < K; j++) put(ver[j]);
    };
    for (int i = 0; i < W; i++)
        for (int j = 0; j < W; j++) put_versions(M[i][j], short_row(W));
    for (int i = 0; i < W; i++)
        for (int j = 0; j < W; j++) put_versions(D[i][j], short_row(W));
    for (int j = 0; j < W; j++) put(c4[j]);
    for (int q = 0; q < kPartial; q++) {
        put(e[q]);
        put_versions(sp[q].dd, short_dot(W));
        for (int i = 0; i < t; i++) put_versions(sp[q].chat[i], short_dot(W));
        for (int i = 0; i < t; i++) put_versions(sp[q].b[i], short_b(W));
    }
    return out.size() == table_entries(W) * 4;
}


// ------------------------------------------------------------------------------------------------
// Canonical-form schedule ("ccf", algo 2).  After the constant pushing above, the partial rounds are a
// TIME-INVARIANT linear system driven by the S-box output s:   w' = A w + b s ,  x' = c^T w + d s
// (A, b, c, d = blocks of the MDS matrix).  In the basis w~ = P w that puts (A, b) in controller canonical
// form, A~ = P A P^-1 is a pure shift except for its last row alpha, and b~ = e_t.  A partial round is then
//      w~_i <- w~_{i+1} (i < t: register renaming) ,  w~_t <- alpha . w~ + s ,  x <- c~ . w~ + d s
// i.e. TWO dot products with two reductions instead of W, and no per-word update loop.  P is folded into the
// full round before (PRE = T M, T = diag(P, 1)); after the last partial round the words are mapped back with
// one dense (W-1)x(W-1) multiply by P^-1.
//
// Diagonal gauge.  x -> x^5 commutes with scaling up to a constant: (lambda x)^5 = lambda^5 x^5.  The kernel
// therefore keeps every word multiplied by a host-chosen non-zero scalar (the "gauge", different per word and
// per round), and the scalars are chosen so that as many matrix entries as possible become exactly 1:
//   * full rounds 0..6:  z^' = G' M G^-5 S(z^ + G c)  with G' chosen so that COLUMN 0 of G' M G^-5 is all
//     ones: every output row is  s_0 + (W-1)-term dot  instead of a W-term dot (W products fewer per round);
//     the last full round has to land on the true state (G' = I) and stays dense;
//   * partial rounds:  x^_q = lambda_q x_q,  w^_{q,i} = nu_{q+i} w~_{q,i}  (a function of q+i, so the shift
//     stays a pure renaming) with  nu_{q+1+t} = lambda_q^5  and  lambda_{q+1} = lambda_q^5 / d : the S-box
//     output then enters BOTH new words with coefficient 1,
//         w^_t' = alpha_q . w^ + s^ ,   x^' = c_q . w^ + s^        (2W-2 multiplications per partial round),
//     at the price of per-round constants alpha_q, c_q (59 x (2W-1) entries with e_q);
//   * the P^-1 stage is followed by a full round, so its rows are gauged too (column 0 all ones).
// Every transformation is exact in F_p and the last round removes the gauge, so outputs are bit-identical.
//
// Table layout (entries of 4 u64): ARK (8W, gauged) | 7 matrices of W x (W-1) (rows without column 0) |
// last matrix W x W | C4' (W) | 59 x { e_q, alpha_q[W-1], c_q[W-1] } | P^-1 gauged, (W-1) x (W-2).
inline size_t ccf_table_entries(int W) {
    const size_t t = W - 1;
    return (size_t)kFull * W + (size_t)(kFull - 1) * W * t + (size_t)W * W + W + (size_t)kPartial * (2 * t + 1) + t * (t - 1);
}

inline bool derive_tables_ccf(int W, const uint64_t* ark, const uint64_t* mds, std::vector<uint64_t>& out) {
    const int t = W - 1;
    auto getF = [](const uint64_t* p) { F f; for (int i = 0; i < 4; i++) f.l[i] = p[i]; return f; };
    Mat M(W, Vec(W));
    for (int i = 0; i < W; i++)
        for (int j = 0; j < W; j++) M[i][j] = getF(mds + (size_t)(i * W + j) * 4);
    std::vector<Vec> c(kFull + kPartial, Vec(W));
    for (int r = 0; r < kFull + kPartial; r++)
        for (int j = 0; j < W; j++) c[r][j] = getF(ark + (size_t)(r * W + j) * 4);
    Mat A(t, Vec(t));
    Vec b(t);
    for (int i = 0; i < t; i++) {
        for (int j = 0; j < t; j++) A[i][j] = M[i][j];
        b[i] = M[i][t];
    }
    // Krylov vectors b, Ab, ..., A^t b;  A^t b = sum_j coef_j A^j b  (Cayley-Hamilton)
    std::vector<Vec> kry(1, b);
    for (int i = 0; i < t; i++) kry.push_back(matvec(A, kry.back()));
    Mat K(t, Vec(t)), Kinv;
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) K[i][j] = kry[j][i];
    if (!invert(K, Kinv)) return false;  // (A, b) not controllable
    Vec alpha = matvec(Kinv, kry[t]);
    // columns v_1..v_t of P^-1:  v_t = b,  v_{j-1} = A v_j - alpha_j v_t
    std::vector<Vec> v(t + 1);
    v[t] = b;
    for (int j = t; j > 1; j--) {
        Vec Av = matvec(A, v[j]);
        v[j - 1].resize(t);
        for (int i = 0; i < t; i++) v[j - 1][i] = sub(Av[i], mul(alpha[j - 1], v[t][i]));
    }
    Mat Pinv(t, Vec(t)), Pm;
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) Pinv[i][j] = v[j + 1][i];
    if (!invert(Pinv, Pm)) return false;
    Mat T(W, Vec(W, kZero)), Tinv(W, Vec(W, kZero));
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) { T[i][j] = Pm[i][j]; Tinv[i][j] = Pinv[i][j]; }
    T[t][t] = kOne;
    Tinv[t][t] = kOne;
    Mat Mt = matmul(matmul(T, M), Tinv);
    // structure check: shift rows, unit input column
    for (int i = 0; i + 1 < t; i++)
        for (int j = 0; j < W; j++) {
            const F& want = (j == i + 1) ? kOne : kZero;
            if (memcmp(Mt[i][j].l, want.l, 32) != 0) return false;
        }
    if (memcmp(Mt[t - 1][t].l, kOne.l, 32) != 0) return false;
    // constant pushing in the transformed coordinates
    std::vector<Vec> ct(c.size());
    for (size_t r = 0; r < c.size(); r++) ct[r] = matvec(T, c[r]);
    Vec d(W, kZero), e(kPartial, kZero);
    for (int q = 0; q + 1 < kPartial; q++) {
        Vec u = matvec(Mt, d);
        for (int j = 0; j < W; j++) u[j] = add(u[j], ct[kHalf + q + 1][j]);
        e[q + 1] = u[t];
        d = u;
        d[t] = kZero;
    }
    Vec tail = matvec(Tinv, matvec(Mt, d));
    Mat pre = matmul(T, M);
    // ---- diagonal gauge -------------------------------------------------------------------------------
    auto pow5 = [](const F& x) { F x2 = mul(x, x); return mul(mul(x2, x2), x); };
    bool ok = true;
    // B = base * diag(g)^-5, then rows scaled so that column 0 is 1 (unit == true); g <- the row scalings
    auto gauge_matrix = [&](const Mat& base, Vec& g, bool unit) {
        const size_t n = base.size(), m = base[0].size();
        Mat B(n, Vec(m));
        for (size_t j = 0; j < m; j++) {
            F s = inv(pow5(g[j]));
            for (size_t i = 0; i < n; i++) B[i][j] = mul(base[i][j], s);
        }
        Vec gn(n, kOne);
        if (unit)
            for (size_t i = 0; i < n; i++) {
                if (is_zero(B[i][0])) { ok = false; return B; }
                gn[i] = inv(B[i][0]);
                for (size_t j = 0; j < m; j++) B[i][j] = mul(B[i][j], gn[i]);
            }
        g = gn;
        return B;
    };
    std::vector<Vec> arkfull(kFull);
    for (int r = 0; r < kHalf; r++) arkfull[r] = c[r];
    arkfull[kHalf] = c[kHalf + kPartial];
    for (int j = 0; j < W; j++) arkfull[kHalf][j] = add(arkfull[kHalf][j], tail[j]);
    for (int r = kHalf + 1; r < kFull; r++) arkfull[r] = c[kPartial + r];
    std::vector<Vec> ark_g(kFull, Vec(W));
    std::vector<Mat> mat_g(kFull);
    Vec g(W, kOne);
    for (int f = 0; f < kHalf; f++) {
        for (int j = 0; j < W; j++) ark_g[f][j] = mul(g[j], arkfull[f][j]);
        mat_g[f] = gauge_matrix(f == kHalf - 1 ? pre : M, g, true);
        if (!ok) return false;
    }
    // g = gauge of (w~_1..w~_t, x) entering the partial rounds
    Vec c4(W);
    for (int j = 0; j < W; j++) c4[j] = mul(g[j], ct[kHalf][j]);
    std::vector<F> nu(kPartial + t + 1, kOne);  // nu[k], k = 1 .. 59 + t
    for (int i = 1; i <= t; i++) nu[i] = g[i - 1];
    F lam = g[t];
    if (is_zero(Mt[t][t])) return false;
    const F dinv = inv(Mt[t][t]);
    std::vector<F> e_g(kPartial);
    std::vector<Vec> alpha_g(kPartial, Vec(t)), crow_g(kPartial, Vec(t));
    for (int q = 0; q < kPartial; q++) {
        e_g[q] = mul(lam, e[q]);
        const F l5 = pow5(lam);
        nu[q + 1 + t] = l5;
        const F lam_next = mul(l5, dinv);
        for (int j = 1; j <= t; j++) {
            const F ninv = inv(nu[q + j]);
            alpha_g[q][j - 1] = mul(mul(l5, Mt[t - 1][j - 1]), ninv);
            crow_g[q][j - 1] = mul(mul(lam_next, Mt[t][j - 1]), ninv);
        }
        lam = lam_next;
    }
    // P^-1 stage: z_i = sum_j Pinv[i][j] w~_j = sum_j (Pinv[i][j] / nu[59 + j + 1]) w^_j ; rows gauged to unit column 0
    Mat pinv_g(t, Vec(t));
    for (int j = 0; j < t; j++) {
        const F ninv = inv(nu[kPartial + j + 1]);
        for (int i = 0; i < t; i++) pinv_g[i][j] = mul(Pinv[i][j], ninv);
    }
    for (int i = 0; i < t; i++) {
        if (is_zero(pinv_g[i][0])) return false;
        g[i] = inv(pinv_g[i][0]);
        for (int j = 0; j < t; j++) pinv_g[i][j] = mul(pinv_g[i][j], g[i]);
    }
    g[t] = lam;
    for (int f = kHalf; f < kFull; f++) {
        for (int j = 0; j < W; j++) ark_g[f][j] = mul(g[j], arkfull[f][j]);
        mat_g[f] = gauge_matrix(M, g, f + 1 < kFull);
        if (!ok) return false;
    }
    out.clear();
    auto put = [&](const F& f) { for (int i = 0; i < 4; i++) out.push_back(f.l[i]); };
    for (int f = 0; f < kFull; f++)
        for (int j = 0; j < W; j++) put(ark_g[f][j]);
    for (int f = 0; f < kFull; f++)
        for (int i = 0; i < W; i++)
            for (int j = (f + 1 < kFull ? 1 : 0); j < W; j++) put(mat_g[f][i][j]);
    for (int j = 0; j < W; j++) put(c4[j]);
    for (int q = 0; q < kPartial; q++) {
        put(e_g[q]);
        for (int j = 0; j < t; j++) put(alpha_g[q][j]);
        for (int j = 0; j < t; j++) put(crow_g[q][j]);
    }
    for (int i = 0; i < t; i++)
        for (int j = 1; j < t; j++) put(pinv_g[i][j]);
    return out.size() == ccf_table_entries(W) * 4;
}

}  // namespace hades_host
