// hades_strategy.hpp -- C++ host-side mirror of the reference crate's public surface for the `perm`
// path, on top of the C ABI (hades_cuda.h).  Header-only.
//
// Reference (paths relative to /root/reference):
//   consts TOTAL_FULL_ROUNDS, PARTIAL_ROUNDS, WIDTH                 src/lib.rs:20-27
//   trait Strategy: perm(&mut self, data: &mut [T]), rounds()        src/strategies.rs:31,140-162
//   ScalarStrategy::new()                                             src/strategies/scalar.rs:12-20
// `CudaStrategy` is the added device strategy: same `perm` contract (in place, exactly WIDTH scalars)
// plus `perm_batch` over `[[BlsScalar; WIDTH]]`.  No CPU fallback: the constructor throws when no
// CUDA device is usable.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "hades_constants.h"
#include "hades_cuda.h"

namespace hades252 {

constexpr std::size_t TOTAL_FULL_ROUNDS = HADES_TOTAL_FULL_ROUNDS;  // lib.rs:22
constexpr std::size_t PARTIAL_ROUNDS = HADES_PARTIAL_ROUNDS;        // lib.rs:26
constexpr std::size_t WIDTH = HADES_WIDTH;                          // lib.rs:27

// Same bytes as dusk_bls12_381::BlsScalar: 4 LE u64 Montgomery limbs, fully reduced.
struct BlsScalar {
    std::uint64_t limbs[4];
    bool operator==(const BlsScalar& o) const {
        return limbs[0] == o.limbs[0] && limbs[1] == o.limbs[1] && limbs[2] == o.limbs[2] && limbs[3] == o.limbs[3];
    }
    bool operator!=(const BlsScalar& o) const { return !(*this == o); }
};
static_assert(sizeof(BlsScalar) == 32 && alignof(BlsScalar) == 8, "layout the FFI relies on");
using State = std::array<BlsScalar, WIDTH>;

struct HadesError : std::runtime_error {
    int status;
    HadesError(int s, const std::string& m) : std::runtime_error("hades status " + std::to_string(s) + ": " + m), status(s) {}
};

// strategies.rs:31 -- algorithm interface; implementors provide `perm`.
template <class T>
struct Strategy {
    virtual ~Strategy() = default;
    virtual void perm(T* data, std::size_t len) = 0;  // strategies.rs:140
    static constexpr std::size_t rounds() { return TOTAL_FULL_ROUNDS + PARTIAL_ROUNDS; }  // strategies.rs:160-162
};

class CudaStrategy : public Strategy<BlsScalar> {
public:
    explicit CudaStrategy(const std::vector<int>& devices = {0}) {
        int rc = hades_init(&ctx_, devices.data(), (int)devices.size(), (std::uint32_t)WIDTH, &HADES_ROUND_CONSTANTS[0][0],
                            HADES_N_ROUND_CONSTANTS, &HADES_MDS_MATRIX_5[0][0]);
        if (rc) throw HadesError(rc, hades_last_error(nullptr));
    }
    CudaStrategy(const CudaStrategy&) = delete;
    CudaStrategy& operator=(const CudaStrategy&) = delete;
    ~CudaStrategy() override { hades_destroy(ctx_); }

    // `Strategy::perm` on one state.  A length other than WIDTH is a programmer error, as in the
    // reference (scalar.rs:48 `copy_from_slice` panics).
    void perm(BlsScalar* data, std::size_t len) override {
        if (len != WIDTH) throw std::invalid_argument("Hades252 perm needs exactly WIDTH scalars");
        check(hades_perm_batch(ctx_, &data[0].limbs[0], 1));
    }
    // `perm_batch(&mut [[BlsScalar; WIDTH]])`
    void perm_batch(State* states, std::size_t n) { check(hades_perm_batch(ctx_, n ? &states[0][0].limbs[0] : nullptr, n)); }
    void perm_batch(std::vector<State>& states) { perm_batch(states.data(), states.size()); }

    BlsScalar merkle_root(const std::vector<BlsScalar>& leaves) {
        BlsScalar root{};
        check(hades_merkle_root(ctx_, leaves.empty() ? nullptr : &leaves[0].limbs[0], leaves.size(), root.limbs));
        return root;
    }
    // ragged tree: any number of leaves, partial nodes hashed under the bitmask of their present children
    BlsScalar merkle_root_ragged(const std::vector<BlsScalar>& leaves) {
        BlsScalar root{};
        check(hades_merkle_root_ragged(ctx_, leaves.empty() ? nullptr : &leaves[0].limbs[0], leaves.size(), root.limbs));
        return root;
    }
    std::vector<BlsScalar> sponge_batch(const std::vector<BlsScalar>& elems, const std::vector<std::uint64_t>& offsets) {
        if (offsets.empty()) throw std::invalid_argument("offsets needs n + 1 entries");
        std::vector<BlsScalar> out(offsets.size() - 1);
        check(hades_sponge_batch(ctx_, elems.empty() ? nullptr : &elems[0].limbs[0], offsets.data(), out.size(),
                                 out.empty() ? nullptr : &out[0].limbs[0]));
        return out;
    }
    // sponge with domain separation: the capacity word starts as `domain` (include/hades_cuda.h)
    std::vector<BlsScalar> sponge_batch(const BlsScalar& domain, const std::vector<BlsScalar>& elems,
                                        const std::vector<std::uint64_t>& offsets) {
        if (offsets.empty() || offsets.front() != 0 || offsets.back() > elems.size())
            throw std::invalid_argument("offsets needs n + 1 entries, offsets[0] == 0 and offsets[n] <= elems.size()");
        std::vector<BlsScalar> out(offsets.size() - 1);
        check(hades_sponge_batch_ds(ctx_, elems.empty() ? nullptr : &elems[0].limbs[0], offsets.data(), out.size(), domain.limbs,
                                    out.empty() ? nullptr : &out[0].limbs[0]));
        return out;
    }
    // batches of at most `max_states` states take the cooperative low-latency kernel (0 disables; default 4736)
    void set_coop_threshold(std::size_t max_states) { check(hades_set_coop_threshold(ctx_, max_states)); }
    // ... and of at most `max_states` states the warp-per-state version of it (0 disables; default 592)
    void set_coop_wide_threshold(std::size_t max_states) { check(hades_set_coop_wide_threshold(ctx_, max_states)); }
    std::string collective() const { return hades_collective(ctx_); }

    hades_ctx* raw() { return ctx_; }

private:
    void check(int rc) {
        if (rc) throw HadesError(rc, hades_last_error(ctx_));
    }
    hades_ctx* ctx_ = nullptr;
};

}  // namespace hades252
