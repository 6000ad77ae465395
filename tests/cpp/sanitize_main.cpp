// Small driver for compute-sanitizer (memcheck / racecheck / synccheck): exercises every kernel family through
// the C ABI on ragged sizes (tails of warps and blocks), Merkle levels and mixed-length sponge messages.
// Not a parity test (that is tests/test_gpu_parity.py): exit code 0 = no CUDA error reported by the library.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/hades_constants.h"
#include "../../include/hades_cuda.h"

static uint64_t rng = 88172645463325252ULL;
static uint64_t next() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }
static void fill(std::vector<uint64_t>& v) {
    for (size_t i = 0; i < v.size(); i++) v[i] = (i % 4 == 3) ? (next() & 0x3fffffffffffffffULL) : next();
}
#define CHECK(x) do { int rc_ = (x); if (rc_) { std::printf("%s -> %d: %s\n", #x, rc_, hades_last_error(ctx)); return 1; } } while (0)

int main() {
    hades_ctx* ctx = nullptr;
    int dev = 0;
    if (hades_init(&ctx, &dev, 1, 5, &HADES_ROUND_CONSTANTS[0][0], HADES_N_ROUND_CONSTANTS, &HADES_MDS_MATRIX_5[0][0])) {
        std::printf("init: %s\n", hades_last_error(nullptr));
        return 77;
    }
    for (int algo = 0; algo < 3; algo++)
        for (int regs : {0, 6, 106, 206}) {   // 106: shape 6 with the cooperative small-batch kernels switched off,
                                              // 206: with the 8-lane kernels only (no warp-per-state kernel)
            if (algo == 0 && regs != 0) continue;
            if (regs >= 100 && algo != 2) continue;
            CHECK(hades_set_variant(ctx, algo, regs % 100));
            if (algo == 2) {
                CHECK(hades_set_coop_threshold(ctx, regs == 106 ? 0 : 4736));
                CHECK(hades_set_coop_wide_threshold(ctx, regs == 206 ? 0 : 592));
            }
            for (size_t n : {1, 31, 33, 127, 129, 300}) {
                std::vector<uint64_t> s(n * 20);
                fill(s);
                CHECK(hades_perm_batch(ctx, s.data(), n));
            }
            std::vector<uint64_t> leaves(256 * 4);
            fill(leaves);
            uint64_t root[4];
            CHECK(hades_merkle_root(ctx, leaves.data(), 256, root));
            for (size_t nl : {2, 5, 130, 255}) CHECK(hades_merkle_root_ragged(ctx, leaves.data(), nl, root));
            std::vector<uint64_t> offsets(201);
            offsets[0] = 0;
            for (int m = 0; m < 200; m++) offsets[m + 1] = offsets[m] + (next() % 11);
            std::vector<uint64_t> elems((offsets[200] + 1) * 4), out(200 * 4);
            fill(elems);
            CHECK(hades_sponge_batch(ctx, elems.data(), offsets.data(), 200, out.data()));
            uint64_t tag[4] = {7, 0, 0, 0};
            CHECK(hades_sponge_batch_ds(ctx, elems.data(), offsets.data(), 200, tag, out.data()));
        }
    {   // pageable memory through the pinned bounce buffers (forced: the batch is far below one chunk)
        CHECK(hades_set_host_path(ctx, 1));
        std::vector<uint64_t> s(1000 * 20);
        fill(s);
        CHECK(hades_perm_batch(ctx, s.data(), 1000));
        CHECK(hades_copy_probe(ctx, s.data(), 1000));
        CHECK(hades_set_host_path(ctx, 0));
    }
    hades_destroy(ctx);
    // other widths: tuned 3 / 9 and per-width dense kernels (7, 4, 12)
    for (uint32_t w : {3u, 9u, 7u, 4u, 12u}) {
        const uint64_t* mds = w == 3 ? &HADES_MDS_MATRIX_3[0][0] : w == 9 ? &HADES_MDS_MATRIX_9[0][0] : nullptr;
        std::vector<uint64_t> fake;
        if (!mds) {  // any canonical values do for a memory check
            fake.resize(w * w * 4);
            fill(fake);
            mds = fake.data();
        }
        if (hades_init(&ctx, &dev, 1, w, &HADES_ROUND_CONSTANTS[0][0], HADES_N_ROUND_CONSTANTS, mds)) {
            std::printf("init w=%u: %s\n", w, hades_last_error(nullptr));
            return 1;
        }
        for (size_t n : {1, 33, 200}) {
            std::vector<uint64_t> s(n * w * 4);
            fill(s);
            CHECK(hades_perm_batch(ctx, s.data(), n));
        }
        hades_destroy(ctx);
    }
    std::puts("sanitize driver OK");
    return 0;
}
