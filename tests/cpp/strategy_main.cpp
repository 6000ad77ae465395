// C++ caller of the drop-in, written like the reference's own tests (scalar.rs:62-74 hades_det,
// README.md:50-65) plus a known-answer check against tests/golden.  Exit code 0 = all good;
// 77 = no CUDA device (the engine refuses to run, which is the expected CPU-box behaviour).
#include <cstdio>
#include <cstring>

#include "../../include/hades_strategy.hpp"

using namespace hades252;

// Montgomery limbs of 17, 19 and of perm([17;5])[0] (tests/golden/hades252_kat.json, "hades_det_17")
static BlsScalar from_u64_mont(const std::uint64_t (&l)[4]) { return BlsScalar{{l[0], l[1], l[2], l[3]}}; }

int main(int argc, char** argv) {
    if (argc != 13) { std::fprintf(stderr, "usage: strategy_main <17 limbs x4> <19 limbs x4> <expected word0 limbs x4>\n"); return 2; }
    std::uint64_t v[12];
    for (int i = 0; i < 12; i++) v[i] = std::strtoull(argv[1 + i], nullptr, 16);
    BlsScalar s17{{v[0], v[1], v[2], v[3]}}, s19{{v[4], v[5], v[6], v[7]}}, want{{v[8], v[9], v[10], v[11]}};
    try {
        CudaStrategy strategy({0});
        State x, y, z;
        x.fill(s17); y.fill(s17); z.fill(s19);
        strategy.perm(x.data(), x.size());
        strategy.perm(y.data(), y.size());
        strategy.perm(z.data(), z.size());
        if (!(x == y) || x == z) { std::puts("hades_det failed"); return 1; }
        if (x[0] != want) { std::puts("known answer mismatch"); return 1; }
        std::vector<State> batch(1000);
        for (auto& st : batch) st.fill(s17);
        strategy.perm_batch(batch);
        for (auto& st : batch) if (!(st == x)) { std::puts("perm_batch mismatch"); return 1; }
        bool threw = false;
        try { strategy.perm(x.data(), 4); } catch (const std::invalid_argument&) { threw = true; }
        if (!threw) { std::puts("wrong length accepted"); return 1; }
        // the lone perm above ran the cooperative kernel; the one-thread kernel must give the same bits
        strategy.set_coop_threshold(0);
        State x2;
        x2.fill(s17);
        strategy.perm(x2.data(), x2.size());
        if (!(x2 == x)) { std::puts("cooperative and one-thread kernels differ"); return 1; }
        // sponge with a zero domain tag equals the plain sponge; a non-zero tag separates the domain
        std::vector<BlsScalar> msg = {s17, s19, s17};
        std::vector<std::uint64_t> off = {0, 3};
        BlsScalar zero{{0, 0, 0, 0}};
        auto plain = strategy.sponge_batch(msg, off), tagged0 = strategy.sponge_batch(zero, msg, off), tagged = strategy.sponge_batch(s19, msg, off);
        if (!(plain[0] == tagged0[0]) || plain[0] == tagged[0]) { std::puts("sponge domain separation failed"); return 1; }
        static_assert(Strategy<BlsScalar>::rounds() == 67, "rounds");
        std::puts("cpp strategy OK");
        return 0;
    } catch (const HadesError& e) {
        std::printf("%s\n", e.what());
        return e.status == HADES_ERR_NO_DEVICE ? 77 : 1;
    }
}
