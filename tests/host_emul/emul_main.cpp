// Host emulation of the device arithmetic (hades252_b200/csrc/*.cuh compiled with
// -DHADES_HOST_EMUL) checked against the CPU oracle (oracle/hades_cpu.c).  No GPU needed.
// Usage: emul_main <ark+mds table file>   (tables as Montgomery u64 limbs: 960*4 then, per
// width in {3,5,9}, W*W*4) ; exits 0 when every comparison is bit-exact.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <pthread.h>
#include <thread>
#include "../../hades252_b200/csrc/hades.cuh"
#include "../../hades252_b200/csrc/coop.cuh"
#include "../../hades252_b200/csrc/host_tables.hpp"

extern "C" {
void oracle_fr_mul(const uint64_t*, const uint64_t*, uint64_t*);
void oracle_fr_add(const uint64_t*, const uint64_t*, uint64_t*);
int oracle_perm(uint64_t* state, int width, const uint64_t* ark, const uint64_t* mds);
}

static std::vector<uint64_t> g_ark, g_mds[10];
template <int W>
struct HostConsts {
    static uint32_t ark(int idx, int k) { return (uint32_t)(g_ark[idx * 4 + k / 2] >> (32 * (k & 1))); }
    static uint32_t mds(int r, int c, int k) { return (uint32_t)(g_mds[W][(r * W + c) * 4 + k / 2] >> (32 * (k & 1))); }
};

static std::vector<uint64_t> g_opt[10];
template <int W>
struct HostTab {
    static uint32_t tab(int entry, int k) { return (uint32_t)(g_opt[W][(size_t)entry * 4 + k / 2] >> (32 * (k & 1))); }
    static const uint32_t* ptr(int entry) { return reinterpret_cast<const uint32_t*>(g_opt[W].data()) + (size_t)entry * 8; }
};

static std::vector<uint64_t> g_ccf[10];
template <int W>
struct HostTabCcf {
    static uint32_t tab(int entry, int k) { return (uint32_t)(g_ccf[W][(size_t)entry * 4 + k / 2] >> (32 * (k & 1))); }
    static const uint32_t* ptr(int entry) { return reinterpret_cast<const uint32_t*>(g_ccf[W].data()) + (size_t)entry * 8; }
};

static uint64_t rng_state = 0x1234567;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static const uint64_t P64[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static void rand_fr(uint64_t* x, int mode) {
    for (int i = 0; i < 4; i++) x[i] = rnd();
    x[3] &= 0x3fffffffffffffffULL;
    if (mode == 1) { memcpy(x, P64, 32); x[0] -= 1 + (rnd() & 3); }           // just below p
    if (mode == 2) { memset(x, 0, 32); x[0] = rnd() & 3; }                     // tiny
    if (mode == 3) { for (int i = 0; i < 4; i++) x[i] = ~0ULL; x[3] = 0x73eda753299d7d47ULL; }  // ff limbs < p
    if (mode == 4) { x[0] = 0; x[1] &= 0xffffffff00000000ULL; }              // zero low limbs
}
static void to32(hades::Fr& f, const uint64_t* x) { for (int k = 0; k < 8; k++) f.l[k] = (uint32_t)(x[k / 2] >> (32 * (k & 1))); }
static void to64(uint64_t* x, const hades::Fr& f) { for (int i = 0; i < 4; i++) x[i] = (uint64_t)f.l[2 * i] | ((uint64_t)f.l[2 * i + 1] << 32); }

template <int W>
static int check_perm(int iters) {
    for (int it = 0; it < iters; it++) {
        uint64_t st[W * 4];
        for (int j = 0; j < W; j++) rand_fr(st + 4 * j, it < 40 ? (it + j) % 5 : 0);
        hades::Fr s[W];
        for (int j = 0; j < W; j++) to32(s[j], st + 4 * j);
        hades::Fr s2[W], s3[W];
        for (int j = 0; j < W; j++) s3[j] = s2[j] = s[j];
        hades::hades_perm_ccf<W, HostTabCcf<W>>(s3);
        hades::hades_perm<W, HostConsts<W>>(s);
        hades::hades_perm_opt<W, HostTab<W>>(s2);
        oracle_perm(st, W, g_ark.data(), g_mds[W].data());
        for (int j = 0; j < W; j++) {
            uint64_t got[4]; to64(got, s[j]);
            if (memcmp(got, st + 4 * j, 32)) { printf("perm W=%d mismatch iter %d word %d\n", W, it, j); return 1; }
            to64(got, s2[j]);
            if (memcmp(got, st + 4 * j, 32)) { printf("perm_opt W=%d mismatch iter %d word %d\n", W, it, j); return 1; }
            to64(got, s3[j]);
            if (memcmp(got, st + 4 * j, 32)) { printf("perm_ccf W=%d mismatch iter %d word %d\n", W, it, j); return 1; }
        }
    }
    return 0;
}

// ---- cooperative kernels (coop.cuh): G host threads play the G lanes of a group (8, or a whole warp); a shuffle is
// an exchange through a shared array between two barriers
static pthread_barrier_t g_bar;
static uint32_t g_xchg[32][16];
static thread_local int tl_lane;
namespace hades {
int coop_emul_lane() { return tl_lane; }
void coop_emul_exchange(uint32_t* v, int n, int xmask, int abs_src) {
    memcpy(g_xchg[tl_lane], v, 4 * n);
    pthread_barrier_wait(&g_bar);
    const int src = abs_src >= 0 ? abs_src : (tl_lane ^ xmask);
    uint32_t tmp[16];
    memcpy(tmp, g_xchg[src], 4 * n);
    pthread_barrier_wait(&g_bar);
    memcpy(v, tmp, 4 * n);
}
}  // namespace hades

template <int G>
static int check_coop(int iters) {
    pthread_barrier_init(&g_bar, nullptr, G);
    for (int it = 0; it < iters; it++) {
        uint64_t st[20];
        for (int j = 0; j < 5; j++) rand_fr(st + 4 * j, it < 40 ? (it + j) % 5 : 0);
        hades::Fr s0[5], res[G][5];
        for (int j = 0; j < 5; j++) to32(s0[j], st + 4 * j);
        std::thread th[G];
        for (int l = 0; l < G; l++)
            th[l] = std::thread([&, l]() {
                tl_lane = l;
                hades::Fr s[5];
                for (int j = 0; j < 5; j++) s[j] = s0[j];
                hades::hades_perm_coop<HostTabCcf<5>, G>(s, l);
                for (int j = 0; j < 5; j++) res[l][j] = s[j];
            });
        for (int l = 0; l < G; l++) th[l].join();
        oracle_perm(st, 5, g_ark.data(), g_mds[5].data());
        for (int l = 0; l < G; l++)
            for (int j = 0; j < 5; j++) {
                uint64_t got[4]; to64(got, res[l][j]);
                if (memcmp(got, st + 4 * j, 32)) { printf("perm_coop<%d> mismatch iter %d lane %d word %d\n", G, it, l, j); return 1; }
            }
    }
    pthread_barrier_destroy(&g_bar);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    g_ark.resize(960 * 4);
    if (fread(g_ark.data(), 8, 960 * 4, f) != 960 * 4) return 2;
    for (int w : {3, 5, 9}) { g_mds[w].resize(w * w * 4); if (fread(g_mds[w].data(), 8, w * w * 4, f) != (size_t)w * w * 4) return 2; }
    fclose(f);
    for (int w : {3, 5, 9}) {
        if (!hades_host::derive_tables(w, g_ark.data(), g_mds[w].data(), g_opt[w])) { printf("derive_tables(%d) failed\n", w); return 1; }
        if (g_opt[w].size() != (size_t)hades::OptLayout<3>::kEntries * 0 + hades_host::table_entries(w) * 4) return 1;
    }
    for (int w : {3, 5, 9})
        if (!hades_host::derive_tables_ccf(w, g_ark.data(), g_mds[w].data(), g_ccf[w])) { printf("derive_tables_ccf(%d) failed\n", w); return 1; }
    if (hades_host::ccf_table_entries(5) != (size_t)hades::CcfLayout<5>::kEntries || hades_host::ccf_table_entries(9) != (size_t)hades::CcfLayout<9>::kEntries ||
        hades_host::ccf_table_entries(3) != (size_t)hades::CcfLayout<3>::kEntries) { printf("ccf layout mismatch\n"); return 1; }
    if (hades_host::table_entries(5) != (size_t)hades::OptLayout<5>::kEntries || hades_host::table_entries(9) != (size_t)hades::OptLayout<9>::kEntries ||
        hades_host::table_entries(3) != (size_t)hades::OptLayout<3>::kEntries) { printf("layout mismatch\n"); return 1; }
    if (argc > 2) {  // dump the derived tables for the Python cross-check
        FILE* o = fopen(argv[2], "wb");
        for (int w : {3, 5, 9}) fwrite(g_opt[w].data(), 8, g_opt[w].size(), o);
        fclose(o);
    }
    // field ops
    for (int it = 0; it < 200000; it++) {
        uint64_t a[4], b[4], want[4], got[4];
        rand_fr(a, it % 7 < 5 ? it % 5 : 0); rand_fr(b, (it / 5) % 5);
        hades::Fr fa, fb, fc; to32(fa, a); to32(fb, b);
        hades::fr_mul(fc, fa, fb); to64(got, fc); oracle_fr_mul(a, b, want);
        if (memcmp(got, want, 32)) { printf("mul mismatch %d\n", it); return 1; }
        hades::fr_add(fc, fa, fb); to64(got, fc); oracle_fr_add(a, b, want);
        if (memcmp(got, want, 32)) { printf("add mismatch %d\n", it); return 1; }
        // x^5
        hades::Fr fx = fa; hades::fr_sbox(fx); to64(got, fx);
        uint64_t x2[4], x4[4]; oracle_fr_mul(a, a, x2); oracle_fr_mul(x2, x2, x4); oracle_fr_mul(x4, a, want);
        if (memcmp(got, want, 32)) { printf("sbox mismatch %d\n", it); return 1; }
    }
    // squaring with adversarial limb patterns (values need not be < p for the limb-level routine,
    // only < 2^256 with a^2/R + p < 2^256: keep the top limb small)
    for (int it = 0; it < 200000; it++) {
        uint64_t a[4], want[4], got[4];
        for (int i = 0; i < 4; i++) { uint64_t r = rnd(); a[i] = (r & 1) ? ~0ULL : (r & 2) ? 0 : (r & 4) ? 0xffffffff00000000ULL : rnd(); }
        a[3] &= 0x3fffffffffffffffULL;
        if (a[3] > 0x73eda753299d7d47ULL || (a[3] == 0x73eda753299d7d47ULL)) a[3] = 0x73eda753299d7d47ULL;
        // force canonical (< p): p's top limb is 0x73eda753299d7d48, so top limb <= ...47 suffices
        hades::Fr fa, fc; to32(fa, a);
        hades::fr_sqr_lazy(fc, fa);
        uint32_t r9[9]; for (int k = 0; k < 8; k++) r9[k] = fc.l[k]; r9[8] = 0;
        hades::canon<0>(fc, r9);
        to64(got, fc); oracle_fr_mul(a, a, want);
        if (memcmp(got, want, 32)) { printf("sqr mismatch %d\n", it); return 1; }
    }
    if (check_perm<5>(300) || check_perm<3>(100) || check_perm<9>(60)) return 1;
    if (check_coop<8>(80)) return 1;
    if (check_coop<32>(60)) return 1;
    printf("host emulation OK\n");
    return 0;
}
