/*
 * CPU oracle (4x64-bit-limb Montgomery) for the Hades252 permutation.
 * TEST INFRASTRUCTURE ONLY: linked / dlopen'ed by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product never calls it.
 *
 * Restates (paths relative to /root/reference):
 *   Strategy::perm / apply_full_round / apply_partial_round   src/strategies.rs:79-157
 *   ScalarStrategy::{add_round_key,quintic_s_box,mul_matrix}  src/strategies/scalar.rs:23-49
 *   ROUND_CONSTANTS / MDS_MATRIX loaders (from_raw)            src/round_constants.rs:29-48,
 *                                                              src/mds_matrix.rs:18-40
 * and, because the field arithmetic is the un-vendored crate dusk-bls12_381 0.13
 * (Cargo.toml:12), that crate's published algorithm: 4 LE u64 limbs in Montgomery
 * form (R = 2^256), multiplication = 4x4 schoolbook product followed by a 4-step
 * word-wise Montgomery reduction and one conditional subtraction of p; addition =
 * limb add followed by a conditional subtraction; every value stays in [0,p).
 *
 * PARITY UNPINNED (see oracle/hades_ref.py header): no KAT exists in the reference
 * and it cannot be executed here.  This file is validated against the independent
 * Python big-int restatement and SURVEY.md 8(c) vectors in tests/.
 *
 * Parallelism: pthreads over contiguous ranges of independent states; this is the
 * stand-in for "rayon over ScalarStrategy" used as the reported CPU baseline.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr_t;

#define FULL_ROUNDS 8
#define PARTIAL_ROUNDS 59
#define MAX_WIDTH 14

static const uint64_t MODULUS[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL,
                                    0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
/* -p^{-1} mod 2^64 */
static const uint64_t INV = 0xfffffffeffffffffULL;
/* R^2 mod p, R = 2^256 */
static const uint64_t R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL,
                               0x05d314967254398fULL, 0x0748d9d99f59ff11ULL};

/* a + b*c + carry -> (lo, new carry) */
static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t *carry) {
    u128 t = (u128)a + (u128)b * c + *carry;
    *carry = (uint64_t)(t >> 64);
    return (uint64_t)t;
}
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t *carry) {
    u128 t = (u128)a + b + *carry;
    *carry = (uint64_t)(t >> 64);
    return (uint64_t)t;
}
static inline uint64_t sbb(uint64_t a, uint64_t b, uint64_t *borrow) {
    u128 t = (u128)a - b - *borrow;
    *borrow = (uint64_t)(t >> 64) & 1;
    return (uint64_t)t;
}

/* r = a - p if a >= p (a given as 4 limbs + an extra top word) */
static inline void reduce_once(fr_t *r, const uint64_t a[4], uint64_t top) {
    uint64_t bw = 0, d[4];
    for (int i = 0; i < 4; i++) d[i] = sbb(a[i], MODULUS[i], &bw);
    uint64_t under = (top < bw); /* borrow out of the 5-word subtraction */
    for (int i = 0; i < 4; i++) r->l[i] = under ? a[i] : d[i];
}

static inline void fr_add(fr_t *r, const fr_t *a, const fr_t *b) {
    uint64_t c = 0, s[4];
    for (int i = 0; i < 4; i++) s[i] = adc(a->l[i], b->l[i], &c);
    reduce_once(r, s, c);
}

/* Fully unrolled on purpose (about 2x faster than the loop form with gcc 13): 16 mac steps for the
 * 4x4 schoolbook product, then the 4-step word-wise Montgomery reduction (k_i = r_i * INV;
 * r += k_i * p << 64 i) and one conditional subtraction.  Same structure as the published
 * `Scalar::mul` / `montgomery_reduce` of the bls12_381 crates. */
#define MAC(a, b, c, carry) ({ u128 t_ = (u128)(a) + (u128)(b) * (c) + (carry); (carry) = (uint64_t)(t_ >> 64); (uint64_t)t_; })
#define ADC(a, b, carry) ({ u128 t_ = (u128)(a) + (b) + (carry); (carry) = (uint64_t)(t_ >> 64); (uint64_t)t_; })
static inline void fr_mul(fr_t *r, const fr_t *x, const fr_t *y) {
    const uint64_t *a = x->l, *b = y->l;
    uint64_t c, r0, r1, r2, r3, r4, r5, r6, r7;
    c = 0; r0 = MAC(0, a[0], b[0], c); r1 = MAC(0, a[0], b[1], c); r2 = MAC(0, a[0], b[2], c); r3 = MAC(0, a[0], b[3], c); r4 = c;
    c = 0; r1 = MAC(r1, a[1], b[0], c); r2 = MAC(r2, a[1], b[1], c); r3 = MAC(r3, a[1], b[2], c); r4 = MAC(r4, a[1], b[3], c); r5 = c;
    c = 0; r2 = MAC(r2, a[2], b[0], c); r3 = MAC(r3, a[2], b[1], c); r4 = MAC(r4, a[2], b[2], c); r5 = MAC(r5, a[2], b[3], c); r6 = c;
    c = 0; r3 = MAC(r3, a[3], b[0], c); r4 = MAC(r4, a[3], b[1], c); r5 = MAC(r5, a[3], b[2], c); r6 = MAC(r6, a[3], b[3], c); r7 = c;
    uint64_t k, c2 = 0;
    k = r0 * INV; c = 0; (void)MAC(r0, k, MODULUS[0], c); r1 = MAC(r1, k, MODULUS[1], c); r2 = MAC(r2, k, MODULUS[2], c); r3 = MAC(r3, k, MODULUS[3], c); r4 = ADC(r4, c, c2);
    k = r1 * INV; c = 0; (void)MAC(r1, k, MODULUS[0], c); r2 = MAC(r2, k, MODULUS[1], c); r3 = MAC(r3, k, MODULUS[2], c); r4 = MAC(r4, k, MODULUS[3], c); r5 = ADC(r5, c, c2);
    k = r2 * INV; c = 0; (void)MAC(r2, k, MODULUS[0], c); r3 = MAC(r3, k, MODULUS[1], c); r4 = MAC(r4, k, MODULUS[2], c); r5 = MAC(r5, k, MODULUS[3], c); r6 = ADC(r6, c, c2);
    k = r3 * INV; c = 0; (void)MAC(r3, k, MODULUS[0], c); r4 = MAC(r4, k, MODULUS[1], c); r5 = MAC(r5, k, MODULUS[2], c); r6 = MAC(r6, k, MODULUS[3], c); r7 = ADC(r7, c, c2);
    const uint64_t hi[4] = {r4, r5, r6, r7};
    reduce_once(r, hi, c2);
}

static inline void fr_square(fr_t *r, const fr_t *a) { fr_mul(r, a, a); }

/* --------------------------------------------------------------- exported Fr ops */
void oracle_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *r) {
    fr_t x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32);
    fr_mul(&z, &x, &y); memcpy(r, &z, 32);
}
void oracle_fr_add(const uint64_t *a, const uint64_t *b, uint64_t *r) {
    fr_t x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32);
    fr_add(&z, &x, &y); memcpy(r, &z, 32);
}
/* BlsScalar::from_raw: canonical integer limbs -> Montgomery limbs (multiply by R^2). */
void oracle_from_raw(const uint64_t *raw, uint64_t *out) {
    fr_t x, y, z; memcpy(&x, raw, 32); memcpy(&y, R2, 32);
    fr_mul(&z, &x, &y); memcpy(out, &z, 32);
}
/* round_constants.rs:29-48 / mds_matrix.rs:18-40 loader: n entries of 32 LE bytes. */
void oracle_load_table(const uint8_t *bytes, size_t n, uint64_t *limbs_out) {
    for (size_t k = 0; k < n; k++) {
        uint64_t raw[4];
        for (int i = 0; i < 4; i++) { /* lib.rs:33-44 u64_from_buffer */
            uint64_t v = 0;
            for (int b = 7; b >= 0; b--) v = (v << 8) | bytes[32 * k + 8 * i + b];
            raw[i] = v;
        }
        oracle_from_raw(raw, limbs_out + 4 * k);
    }
}

/* -------------------------------------------------------------------- permutation */
static inline void quintic_s_box(fr_t *x) { /* scalar.rs:32-34 */
    fr_t x2, x4;
    fr_square(&x2, x);
    fr_square(&x4, &x2);
    fr_mul(x, &x4, x);
}

static inline void mul_matrix(fr_t *v, int w, const fr_t *mds) { /* scalar.rs:36-49 */
    fr_t res[MAX_WIDTH];
    memset(res, 0, sizeof(fr_t) * w);
    for (int j = 0; j < w; j++)
        for (int k = 0; k < w; k++) {
            fr_t t;
            fr_mul(&t, &mds[k * w + j], &v[j]);
            fr_add(&res[k], &res[k], &t);
        }
    memcpy(v, res, sizeof(fr_t) * w);
}

static void perm_one(fr_t *s, int w, const fr_t *ark, const fr_t *mds) {
    const fr_t *c = ark; /* strategies.rs:141 */
    for (int r = 0; r < FULL_ROUNDS + PARTIAL_ROUNDS; r++) {
        int full = (r < FULL_ROUNDS / 2) || (r >= FULL_ROUNDS / 2 + PARTIAL_ROUNDS);
        for (int i = 0; i < w; i++) fr_add(&s[i], &s[i], c++); /* scalar.rs:23-30 */
        if (full) for (int i = 0; i < w; i++) quintic_s_box(&s[i]); /* strategies.rs:115 */
        else quintic_s_box(&s[w - 1]);                              /* strategies.rs:89 */
        mul_matrix(s, w, mds);
    }
}

int oracle_perm(uint64_t *state, int width, const uint64_t *ark, const uint64_t *mds) {
    if (width < 2 || width > MAX_WIDTH) return 1;
    perm_one((fr_t *)state, width, (const fr_t *)ark, (const fr_t *)mds);
    return 0;
}

/* ------------------------------------------------------------------ thread helper */
typedef void (*range_fn)(void *ctx, size_t lo, size_t hi);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = p; j->fn(j->ctx, j->lo, j->hi); return NULL; }

static void parallel_for(size_t n, int nthreads, range_fn fn, void *ctx) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    if (nthreads == 1) { fn(ctx, 0, n); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (job_t){fn, ctx, n * t / nthreads, n * (t + 1) / nthreads};
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* --------------------------------------------------------------------- perm_batch */
typedef struct { uint64_t *states; int w; const fr_t *ark, *mds; } batch_ctx;
static void batch_range(void *p, size_t lo, size_t hi) {
    batch_ctx *c = p;
    for (size_t i = lo; i < hi; i++) perm_one((fr_t *)(c->states + i * 4 * c->w), c->w, c->ark, c->mds);
}
int oracle_perm_batch(uint64_t *states, size_t n, int width, const uint64_t *ark,
                      const uint64_t *mds, int nthreads) {
    if (width < 2 || width > MAX_WIDTH) return 1;
    batch_ctx c = {states, width, (const fr_t *)ark, (const fr_t *)mds};
    parallel_for(n, nthreads, batch_range, &c);
    return 0;
}

/* ------------------------------------------------------------------------- merkle */
typedef struct { const fr_t *in; fr_t *out; const fr_t *ark, *mds; fr_t mask; } merkle_ctx;
static void merkle_range(void *p, size_t lo, size_t hi) {
    merkle_ctx *c = p;
    for (size_t i = lo; i < hi; i++) {
        fr_t s[5];
        s[0] = c->mask;
        memcpy(&s[1], &c->in[4 * i], 4 * sizeof(fr_t));
        perm_one(s, 5, c->ark, c->mds);
        c->out[i] = s[1];
    }
}
/* 4-ary tree, node = perm([15, c0..c3])[1]; n_leaves must be a power of 4 (>= 1). */
int oracle_merkle_root(const uint64_t *leaves, size_t n_leaves, const uint64_t *ark,
                       const uint64_t *mds, uint64_t *root, int nthreads) {
    if (n_leaves == 0 || (n_leaves & (n_leaves - 1))) return 1;
    int lg = 0; while (((size_t)1 << lg) < n_leaves) lg++;
    if (lg & 1) return 1;
    if (n_leaves == 1) { memcpy(root, leaves, 32); return 0; }
    merkle_ctx c; c.ark = (const fr_t *)ark; c.mds = (const fr_t *)mds;
    uint64_t raw15[4] = {15, 0, 0, 0};
    oracle_from_raw(raw15, c.mask.l);
    fr_t *bufA = malloc(sizeof(fr_t) * (n_leaves / 4)), *bufB = malloc(sizeof(fr_t) * (n_leaves / 16 + 1));
    const fr_t *in = (const fr_t *)leaves; fr_t *out = bufA;
    size_t n = n_leaves;
    while (n > 1) {
        c.in = in; c.out = out;
        parallel_for(n / 4, nthreads, merkle_range, &c);
        n /= 4; in = out; out = (out == bufA) ? bufB : bufA;
    }
    memcpy(root, in, 32);
    free(bufA); free(bufB);
    return 0;
}

/* Ragged tree (any n_leaves >= 1), all interior levels written to `tree` (level 1 first, root last;
 * oracle_merkle_tree_nodes(n) nodes): a parent with k present children = perm([2^k - 1, c.., 0..])[1].
 * Restates oracle/hades_ref.py merkle_levels (build-defined convention). */
typedef struct { const fr_t *in; fr_t *out; size_t n_in; const fr_t *ark, *mds; fr_t mask[5]; } rtree_ctx;
static void rtree_range(void *p, size_t lo, size_t hi) {
    rtree_ctx *c = p;
    for (size_t i = lo; i < hi; i++) {
        size_t k = c->n_in - 4 * i; if (k > 4) k = 4;
        fr_t s[5]; memset(s, 0, sizeof s);
        s[0] = c->mask[k];
        memcpy(&s[1], &c->in[4 * i], k * sizeof(fr_t));
        perm_one(s, 5, c->ark, c->mds);
        c->out[i] = s[1];
    }
}
size_t oracle_merkle_tree_nodes(size_t n_leaves) {
    size_t total = 0, m = n_leaves;
    while (m > 1) { m = (m + 3) / 4; total += m; }
    return total;
}
int oracle_merkle_tree(const uint64_t *leaves, size_t n_leaves, const uint64_t *ark, const uint64_t *mds,
                       uint64_t *tree, int nthreads) {
    if (n_leaves == 0) return 1;
    rtree_ctx c; c.ark = (const fr_t *)ark; c.mds = (const fr_t *)mds;
    for (uint64_t k = 1; k <= 4; k++) { uint64_t raw[4] = {(1ULL << k) - 1, 0, 0, 0}; oracle_from_raw(raw, c.mask[k].l); }
    const fr_t *in = (const fr_t *)leaves; fr_t *out = (fr_t *)tree;
    size_t m = n_leaves;
    while (m > 1) {
        size_t n_out = (m + 3) / 4;
        c.in = in; c.out = out; c.n_in = m;
        parallel_for(n_out, nthreads, rtree_range, &c);
        in = out; out += n_out; m = n_out;
    }
    return 0;
}

/* ------------------------------------------------------------------------- sponge */
typedef struct { const fr_t *elems; const uint64_t *off; fr_t *out; const fr_t *ark, *mds; fr_t one; fr_t tag; } sponge_ctx;
static void sponge_range(void *p, size_t lo, size_t hi) {
    sponge_ctx *c = p;
    for (size_t m = lo; m < hi; m++) {
        fr_t s[5]; memset(s, 0, sizeof s);
        s[0] = c->tag;  /* capacity word: zero, or the domain tag (oracle_sponge_batch_ds) */
        size_t b = c->off[m], e = c->off[m + 1];
        int done = 0;
        while (!done) {
            for (int k = 0; k < 4; k++) {
                if (b < e) { fr_add(&s[1 + k], &s[1 + k], &c->elems[b]); b++; }
                else if (!done) { fr_add(&s[1 + k], &s[1 + k], &c->one); done = 1; }
                /* remaining pad elements are zero */
            }
            perm_one(s, 5, c->ark, c->mds);
        }
        c->out[m] = s[1];
    }
}
/* rate 4 / capacity 1, pad = single 1 then zeros; digest = word 1.  CSR offsets (n+1). */
int oracle_sponge_batch(const uint64_t *elems, const uint64_t *offsets, size_t n_msgs,
                        const uint64_t *ark, const uint64_t *mds, uint64_t *out, int nthreads) {
    sponge_ctx c = {(const fr_t *)elems, offsets, (fr_t *)out, (const fr_t *)ark, (const fr_t *)mds, {{0}}, {{0}}};
    uint64_t raw1[4] = {1, 0, 0, 0};
    oracle_from_raw(raw1, c.one.l);
    parallel_for(n_msgs, nthreads, sponge_range, &c);
    return 0;
}
/* same with domain separation: the capacity word (word 0) starts as `tag` (Montgomery limbs, < p) instead of zero
 * (hades_ref.py sponge(message, domain)).  Build-defined like the sponge itself: the reference holds no sponge. */
int oracle_sponge_batch_ds(const uint64_t *elems, const uint64_t *offsets, size_t n_msgs, const uint64_t *tag,
                           const uint64_t *ark, const uint64_t *mds, uint64_t *out, int nthreads) {
    sponge_ctx c = {(const fr_t *)elems, offsets, (fr_t *)out, (const fr_t *)ark, (const fr_t *)mds, {{0}}, {{0}}};
    uint64_t raw1[4] = {1, 0, 0, 0};
    oracle_from_raw(raw1, c.one.l);
    memcpy(c.tag.l, tag, 32);
    parallel_for(n_msgs, nthreads, sponge_range, &c);
    return 0;
}

/* --------------------------------------------------------------- synthetic inputs */
static inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
/* field element number idx (global index), limb l = splitmix64(seed + 4*idx + l),
 * top limb masked to 62 bits (value < 2^254 < p).  SURVEY.md 8(d). */
void oracle_gen_elems(uint64_t *out, uint64_t first_elem, size_t n_elems, uint64_t seed) {
    for (size_t i = 0; i < n_elems; i++)
        for (int l = 0; l < 4; l++) {
            uint64_t v = splitmix64(seed + (first_elem + i) * 4 + l);
            out[4 * i + l] = (l == 3) ? (v & 0x3fffffffffffffffULL) : v;
        }
}
/* 256-bit digest of a limb array: dig[0..1] = xor / wrapping sum of splitmix64(limb ^ splitmix64(global limb index)). */
void oracle_digest(const uint64_t *limbs, uint64_t first_limb, size_t n_limbs, uint64_t dig[4]) {
    uint64_t x = 0, s = 0, x2 = 0, s2 = 0;
    for (size_t i = 0; i < n_limbs; i++) {
        uint64_t h = splitmix64(limbs[i] ^ splitmix64(first_limb + i));
        x ^= h; s += h; x2 ^= limbs[i]; s2 += limbs[i];
    }
    dig[0] = x; dig[1] = s; dig[2] = x2; dig[3] = s2;
}
