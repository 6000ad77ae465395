// libhades_b200.so -- host side of the batched Hades252 engine: context, constant upload,
// multi-device sharding, chunked H2D/compute/D2H pipeline, Merkle and sponge drivers, C ABI
// (include/hades_cuda.h).  No CPU fallback anywhere: every entry point either runs CUDA kernels
// or fails with a status code.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hades_cuda.h"
#include "fr.cuh"
#include "host_tables.hpp"
#include "util_kernels.cuh"
#include "width_ops.hpp"

using namespace hades;

namespace {

constexpr int kRounds = 67;                      // 8 full + 59 partial (src/lib.rs:20-27)
constexpr int kNumBuf = 3;                       // chunk buffers (and streams) per device
constexpr size_t kChunkBytes = (size_t)96 << 20;  // target bytes per pipeline chunk
// Batches / Merkle levels up to this many states run the cooperative 8-lanes-per-state kernels (coop.cuh): below it
// the one-thread-per-state kernel is latency-bound (one warp per scheduler, ~270 us whatever the size).
// Measured (profiles/r02_latency_small_batches.txt): 127 us up to 2368 states (one 16-state block per SM), 178 us up to
// 4736, 333 us at 8192 -- against 269 us for the one-thread kernel at any size up to 16 384.
constexpr int kDefaultCoopMax = 4736;

struct DeviceState {
    int ordinal = 0;
    uint64_t* generic_tables = nullptr;  // widths without a tuned kernel: ark ++ mds in global memory
    cudaStream_t streams[kNumBuf] = {nullptr, nullptr, nullptr};
    uint64_t* chunk[kNumBuf] = {nullptr, nullptr, nullptr};
    size_t chunk_bytes = 0;
};

thread_local std::string g_init_error;

}  // namespace

struct hades_ctx {
    uint32_t width = 0;
    const WidthOps* ops2[3] = {nullptr, nullptr, nullptr};  // [algo]
    const WidthOps* ops() const { return ops2[variant.algo]; }
    bool generic() const { return ops2[0] == nullptr; }  // no tuned kernel for this width
    Variant variant = {1, 0};  // optimised schedule, <=128 registers
    bool has_ccf = false;      // canonical-form tables derived (algo 2 available)
    std::vector<DeviceState> devs;
    mutable std::string err;
    uint64_t launches = 0;
};

namespace {

std::mutex g_tables_mutex;
// (device ordinal, width or 0 for ark) -> bytes resident in __constant__ memory
std::map<std::pair<int, int>, std::vector<uint64_t>> g_tables;

int fail(hades_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_init_error = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                            \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, HADES_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                           \
    } while (0)

bool valid_dev(const hades_ctx* ctx, int dev_index) { return ctx && dev_index >= 0 && dev_index < (int)ctx->devs.size(); }

// Tables resident in __constant__ memory per (device ordinal, width): the reference's tables are
// compile-time constants of the crate, so they are process-wide here as well.
int upload_tables(hades_ctx* ctx, int ordinal, const std::vector<uint64_t>& dense, const std::vector<uint64_t>& opt,
                  const std::vector<uint64_t>& ccf) {
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    auto it = g_tables.find({ordinal, (int)ctx->width});
    if (it != g_tables.end()) {
        if (it->second != dense)
            return fail(ctx, HADES_ERR_CONSTANTS,
                        "constant tables for width %u already resident on device %d with different contents", ctx->width, ordinal);
        return HADES_OK;
    }
    CUDA_TRY(ctx, ctx->ops2[0]->upload(dense.data()));
    CUDA_TRY(ctx, ctx->ops2[1]->upload(opt.data()));
    if (!ccf.empty()) CUDA_TRY(ctx, ctx->ops2[2]->upload(ccf.data()));
    g_tables[{ordinal, (int)ctx->width}] = dense;
    return HADES_OK;
}

int launch_perm_w(hades_ctx* ctx, uint64_t* d_states, size_t n, cudaStream_t stream, const DeviceState* dev = nullptr) {
    if (n == 0) return HADES_OK;
    ctx->launches++;
    if (ctx->generic()) {
        if (!dev) return fail(ctx, HADES_ERR_INVALID_ARG, "internal: generic launch without a device");
        CUDA_TRY(ctx, generic_launch_perm(d_states, n, (int)ctx->width, dev->generic_tables, stream));
        return HADES_OK;
    }
    CUDA_TRY(ctx, ctx->ops()->launch_perm(ctx->variant, d_states, n, stream));
    return HADES_OK;
}

int ensure_chunks(hades_ctx* ctx, DeviceState& d, size_t bytes) {
    if (d.chunk_bytes >= bytes) return HADES_OK;
    for (int b = 0; b < kNumBuf; b++) {
        if (d.chunk[b]) CUDA_TRY(ctx, cudaFree(d.chunk[b]));
        d.chunk[b] = nullptr;
    }
    d.chunk_bytes = 0;
    for (int b = 0; b < kNumBuf; b++) CUDA_TRY(ctx, cudaMalloc(&d.chunk[b], bytes));
    d.chunk_bytes = bytes;
    return HADES_OK;
}

int merkle_reduce(hades_ctx* ctx, const uint64_t* d_nodes, size_t n_nodes, int levels, uint64_t* d_scratch,
                  uint64_t* d_out, cudaStream_t stream) {
    const uint64_t* in = d_nodes;
    uint64_t* bufA = d_scratch;
    uint64_t* bufB = d_scratch + (n_nodes / 4) * 4;
    size_t n = n_nodes;
    for (int l = 0; l < levels; l++) {
        size_t n_out = n / 4;
        uint64_t* out = (l == levels - 1) ? d_out : ((l & 1) ? bufB : bufA);
        ctx->launches++;
        CUDA_TRY(ctx, ctx->ops()->launch_merkle_level(ctx->variant, in, out, n_out, n, stream));
        in = out;
        n = n_out;
    }
    return HADES_OK;
}

int log4_exact(size_t n) {  // k if n == 4^k else -1
    if (n == 0 || (n & (n - 1))) return -1;
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    return (lg & 1) ? -1 : lg / 2;
}

}  // namespace

extern "C" {

int hades_init(hades_ctx** out, const int* devices, int n_dev, uint32_t width, const uint64_t* ark_limbs,
               size_t n_ark, const uint64_t* mds_limbs) {
    if (!out || !ark_limbs || !mds_limbs || n_dev < 1)
        return fail(nullptr, HADES_ERR_INVALID_ARG, "hades_init: null pointer or n_dev < 1");
    *out = nullptr;
    if (width < 2 || width > 14)
        return fail(nullptr, HADES_ERR_INVALID_ARG, "hades_init: width %u out of range (2..14; 67*width round constants must fit 960)", width);
    const bool tuned = width == 3 || width == 5 || width == 9;
    if ((size_t)kRounds * width > n_ark)
        return fail(nullptr, HADES_ERR_OUT_OF_CONSTANTS, "Hades252 out of ARK constants: need %zu, got %zu",
                    (size_t)kRounds * width, n_ark);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1)
        return fail(nullptr, HADES_ERR_NO_DEVICE, "no CUDA device available (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    hades_ctx* ctx = new hades_ctx();
    ctx->width = width;
    if (tuned)
        for (int a = 0; a < 3; a++) ctx->ops2[a] = width == 3 ? width_ops_3(a) : width == 5 ? width_ops_5(a) : width_ops_9(a);
    ctx->variant = Variant{1, width == 9 ? 7 : 6};  // lockstep 128-thread blocks: x7 (W=3), x5 (W=5), x3 (W=9) per SM
    // dense table = ROUND_CONSTANTS[0..67W) ++ MDS_MATRIX; optimised table derived from it (host_tables.hpp)
    std::vector<uint64_t> dense(ark_limbs, ark_limbs + (size_t)kRounds * width * 4), opt, ccf;
    dense.insert(dense.end(), mds_limbs, mds_limbs + (size_t)width * width * 4);
    int rc = HADES_OK;
    if (tuned) {
        if (dense.size() != ctx->ops2[0]->table_u64) rc = fail(nullptr, HADES_ERR_INVALID_ARG, "internal: dense table size");
        if (rc == HADES_OK && (!hades_host::derive_tables((int)width, ark_limbs, mds_limbs, opt) || opt.size() != ctx->ops2[1]->table_u64))
            rc = fail(nullptr, HADES_ERR_CONSTANTS, "could not derive the sparse partial-round tables (singular MDS sub-matrix)");
        // The gauged canonical-form schedule (default) needs a controllable MDS block and non-zero pivots; with
        // constants where that fails (never with the reference's assets) the sparse schedule stays the default.
        if (rc == HADES_OK) {
            ctx->has_ccf = hades_host::derive_tables_ccf((int)width, ark_limbs, mds_limbs, ccf) && ccf.size() == ctx->ops2[2]->table_u64;
            if (!ctx->has_ccf) ccf.clear();
            else {
                ctx->variant.algo = 2;
                if (width == 5) ctx->variant.coop_max = kDefaultCoopMax;
            }
        }
    }
    for (int g = 0; g < n_dev && rc == HADES_OK; g++) {
        DeviceState d;
        d.ordinal = devices ? devices[g] : g;
        if (d.ordinal < 0 || d.ordinal >= count) {
            rc = fail(nullptr, HADES_ERR_NO_DEVICE, "device ordinal %d out of range (0..%d)", d.ordinal, count - 1);
            break;
        }
        ctx->devs.push_back(d);
    }
    for (size_t g = 0; g < ctx->devs.size() && rc == HADES_OK; g++) {
        DeviceState& d = ctx->devs[g];
        auto step = [&]() -> int {
            CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
            for (int b = 0; b < kNumBuf; b++) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&d.streams[b], cudaStreamNonBlocking));
            if (!tuned) {  // generic kernel: per-context tables in global memory
                CUDA_TRY(ctx, generic_upload_modulus());
                CUDA_TRY(ctx, cudaMalloc(&d.generic_tables, dense.size() * 8));
                CUDA_TRY(ctx, cudaMemcpy(d.generic_tables, dense.data(), dense.size() * 8, cudaMemcpyHostToDevice));
                return HADES_OK;
            }
            return upload_tables(ctx, d.ordinal, dense, opt, ccf);
        };
        rc = step();
    }
    if (rc != HADES_OK) {
        g_init_error = ctx->err.empty() ? g_init_error : ctx->err;
        hades_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return HADES_OK;
}

void hades_destroy(hades_ctx* ctx) {
    if (!ctx) return;
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.ordinal);
        for (int b = 0; b < kNumBuf; b++) {
            if (d.streams[b]) cudaStreamSynchronize(d.streams[b]);
            if (d.chunk[b]) cudaFree(d.chunk[b]);
            if (d.streams[b]) cudaStreamDestroy(d.streams[b]);
        }
        if (d.generic_tables) cudaFree(d.generic_tables);
    }
    delete ctx;
}

const char* hades_last_error(const hades_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }
uint32_t hades_width(const hades_ctx* ctx) { return ctx ? ctx->width : 0; }
int hades_device_count(const hades_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
uint64_t hades_launch_count(const hades_ctx* ctx) { return ctx ? ctx->launches : 0; }

int hades_perm_batch_dev(hades_ctx* ctx, int dev_index, uint64_t* d_states, size_t n, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    if (n == 0) return HADES_OK;
    if (!d_states || ((uintptr_t)d_states & 15)) return fail(ctx, HADES_ERR_INVALID_ARG, "d_states must be non-null and 16-byte aligned");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    return launch_perm_w(ctx, d_states, n, (cudaStream_t)stream, &ctx->devs[dev_index]);
}

int hades_perm_batch(hades_ctx* ctx, uint64_t* host_states, size_t n) {
    if (!ctx) return fail(ctx, HADES_ERR_INVALID_ARG, "null context");
    if (n == 0) return HADES_OK;
    if (!host_states) return fail(ctx, HADES_ERR_INVALID_ARG, "null states pointer");
    const size_t state_bytes = (size_t)ctx->width * 32;
    const size_t G = ctx->devs.size();
    size_t chunk_states = std::max<size_t>(kPermThreads, kChunkBytes / state_bytes / kPermThreads * kPermThreads);
    // Medium batches: split into kNumBuf chunks so that H2D, kernel and D2H still overlap.  Small batches
    // (under kMinSplitBytes per chunk) go as ONE chunk: a kernel launch is one ~340 us wave regardless of its
    // size up to ~16k states, and with pageable host memory the copies are synchronous, so splitting would
    // only serialise several such waves.
    constexpr size_t kMinSplitBytes = (size_t)8 << 20;
    size_t per_dev = (n + G - 1) / G;
    if (per_dev < chunk_states * kNumBuf) {
        if (per_dev * state_bytes >= kMinSplitBytes * kNumBuf)
            chunk_states = std::max<size_t>(kPermThreads, (per_dev / kNumBuf + kPermThreads) / kPermThreads * kPermThreads);
        else
            chunk_states = std::max<size_t>(kPermThreads, (per_dev + kPermThreads - 1) / kPermThreads * kPermThreads);
    }
    std::vector<size_t> lo(G), hi(G), next(G);
    size_t max_chunks = 0;
    for (size_t g = 0; g < G; g++) {
        lo[g] = n * g / G;
        hi[g] = n * (g + 1) / G;
        next[g] = lo[g];
        if (hi[g] > lo[g]) {
            CUDA_TRY(ctx, cudaSetDevice(ctx->devs[g].ordinal));
            int r = ensure_chunks(ctx, ctx->devs[g], std::min(chunk_states, hi[g] - lo[g]) * state_bytes);
            if (r) return r;
            max_chunks = std::max(max_chunks, (hi[g] - lo[g] + chunk_states - 1) / chunk_states);
        }
    }
    int rc = HADES_OK;
    for (size_t c = 0; c < max_chunks && rc == HADES_OK; c++) {
        for (size_t g = 0; g < G && rc == HADES_OK; g++) {
            if (next[g] >= hi[g]) continue;
            DeviceState& d = ctx->devs[g];
            size_t cnt = std::min(chunk_states, hi[g] - next[g]);
            int b = (int)(c % kNumBuf);
            uint64_t* h = host_states + next[g] * ctx->width * 4;
            auto step = [&]() -> int {
                CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
                CUDA_TRY(ctx, cudaMemcpyAsync(d.chunk[b], h, cnt * state_bytes, cudaMemcpyHostToDevice, d.streams[b]));
                int r = launch_perm_w(ctx, d.chunk[b], cnt, d.streams[b], &d);
                if (r) return r;
                CUDA_TRY(ctx, cudaMemcpyAsync(h, d.chunk[b], cnt * state_bytes, cudaMemcpyDeviceToHost, d.streams[b]));
                return HADES_OK;
            };
            rc = step();
            next[g] += cnt;
        }
    }
    for (size_t g = 0; g < G; g++) {
        cudaSetDevice(ctx->devs[g].ordinal);
        for (int b = 0; b < kNumBuf; b++) {
            cudaError_t e = cudaStreamSynchronize(ctx->devs[g].streams[b]);
            if (e != cudaSuccess && rc == HADES_OK)
                rc = fail(ctx, HADES_ERR_CUDA, "perm_batch pipeline failed on device %d: %s", ctx->devs[g].ordinal,
                          cudaGetErrorString(e));
        }
    }
    return rc;
}

int hades_merkle_reduce_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_nodes, size_t n_nodes, int levels,
                            uint64_t* d_scratch, uint64_t* d_out, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "merkle needs a width-5 context");
    if (levels < 0 || !d_nodes || !d_out) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer or negative levels");
    if (levels > 31 || (n_nodes >> (2 * levels)) == 0 || (n_nodes & (((size_t)1 << (2 * levels)) - 1)))
        return fail(ctx, HADES_ERR_NOT_POWER_OF_4, "n_nodes=%zu is not a multiple of 4^%d", n_nodes, levels);
    if (levels > 1 && !d_scratch) return fail(ctx, HADES_ERR_INVALID_ARG, "scratch required for more than one level");
    if (((uintptr_t)d_nodes | (uintptr_t)d_out | (uintptr_t)d_scratch) & 15)
        return fail(ctx, HADES_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    if (levels == 0) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_out, d_nodes, n_nodes * 32, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return HADES_OK;
    }
    return merkle_reduce(ctx, d_nodes, n_nodes, levels, d_scratch, d_out, (cudaStream_t)stream);
}

size_t hades_merkle_tree_nodes(size_t n_leaves) {
    size_t total = 0, m = n_leaves;
    while (m > 1) {
        m = (m + 3) / 4;
        total += m;
    }
    return total;
}

int hades_merkle_tree_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_leaves, size_t n_leaves, uint64_t* d_tree,
                          void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "merkle needs a width-5 context");
    if (n_leaves == 0 || !d_leaves) return fail(ctx, HADES_ERR_INVALID_ARG, "a tree needs at least one leaf");
    if (n_leaves == 1) return HADES_OK;  // the leaf is the root; no interior node
    if (!d_tree) return fail(ctx, HADES_ERR_INVALID_ARG, "null tree pointer");
    if (((uintptr_t)d_leaves | (uintptr_t)d_tree) & 15) return fail(ctx, HADES_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    const uint64_t* in = d_leaves;
    uint64_t* out = d_tree;
    for (size_t m = n_leaves; m > 1;) {
        const size_t n_out = (m + 3) / 4;
        ctx->launches++;
        CUDA_TRY(ctx, ctx->ops()->launch_merkle_level(ctx->variant, in, out, n_out, m, (cudaStream_t)stream));
        in = out;
        out += n_out * 4;
        m = n_out;
    }
    return HADES_OK;
}

int hades_merkle_open_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_leaves, const uint64_t* d_tree, size_t n_leaves,
                          const uint64_t* d_index, size_t n_open, uint64_t* d_branch, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    if (n_leaves == 0 || !d_leaves) return fail(ctx, HADES_ERR_INVALID_ARG, "a tree needs at least one leaf");
    int levels = 0;
    for (size_t m = n_leaves; m > 1; m = (m + 3) / 4) levels++;
    if (n_open == 0 || levels == 0) return HADES_OK;  // nothing to write (a single leaf has an empty path)
    if (!d_tree || !d_index || !d_branch) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (((uintptr_t)d_leaves | (uintptr_t)d_tree | (uintptr_t)d_branch) & 15)
        return fail(ctx, HADES_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    const size_t chunks = n_open * (size_t)levels * 8;
    const unsigned blocks = (unsigned)std::min<size_t>((chunks + 255) / 256, 148 * 16);
    merkle_open_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(d_leaves),
                                                                 reinterpret_cast<const uint4*>(d_tree), n_leaves, d_index, n_open,
                                                                 levels, reinterpret_cast<uint4*>(d_branch));
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return HADES_OK;
}

int hades_merkle_root_ragged(hades_ctx* ctx, const uint64_t* host_leaves, size_t n_leaves, uint64_t root[4]) {
    if (!ctx || !host_leaves || !root) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "merkle needs a width-5 context");
    if (n_leaves == 0) return fail(ctx, HADES_ERR_INVALID_ARG, "a tree needs at least one leaf");
    if (n_leaves == 1) {
        memcpy(root, host_leaves, 32);
        return HADES_OK;
    }
    DeviceState& d = ctx->devs[0];
    uint64_t *d_leaves = nullptr, *d_tree = nullptr;
    const size_t nodes = hades_merkle_tree_nodes(n_leaves);
    auto step = [&]() -> int {
        CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
        CUDA_TRY(ctx, cudaMalloc(&d_leaves, n_leaves * 32));
        CUDA_TRY(ctx, cudaMalloc(&d_tree, nodes * 32));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_leaves, host_leaves, n_leaves * 32, cudaMemcpyHostToDevice, d.streams[0]));
        int r = hades_merkle_tree_dev(ctx, 0, d_leaves, n_leaves, d_tree, d.streams[0]);
        if (r) return r;
        CUDA_TRY(ctx, cudaMemcpyAsync(root, d_tree + (nodes - 1) * 4, 32, cudaMemcpyDeviceToHost, d.streams[0]));
        CUDA_TRY(ctx, cudaStreamSynchronize(d.streams[0]));
        return HADES_OK;
    };
    int rc = step();
    cudaFree(d_leaves);
    cudaFree(d_tree);
    return rc;
}

int hades_merkle_root(hades_ctx* ctx, const uint64_t* host_leaves, size_t n_leaves, uint64_t root[4]) {
    if (!ctx || !host_leaves || !root) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "merkle needs a width-5 context");
    int depth = log4_exact(n_leaves);
    if (depth < 0) return fail(ctx, HADES_ERR_NOT_POWER_OF_4, "number of leaves (%zu) must be a power of 4", n_leaves);
    if (depth == 0) {
        memcpy(root, host_leaves, 32);
        return HADES_OK;
    }
    // use the largest power-of-two device count that leaves >= 4 whole subtrees' worth of work each
    size_t G = 1;
    while (G * 2 <= ctx->devs.size() && n_leaves / (G * 2) >= 1024) G *= 2;
    size_t per_dev = n_leaves / G;                 // = 4^a or 2*4^a
    int sub_levels = 0;                            // levels each device can reduce on its own range
    while (((per_dev >> (2 * (sub_levels + 1))) << (2 * (sub_levels + 1))) == per_dev && (per_dev >> (2 * (sub_levels + 1))) >= 1)
        sub_levels++;
    size_t roots_per_dev = per_dev >> (2 * sub_levels);  // 1 or 2
    size_t n_roots = roots_per_dev * G;
    std::vector<uint64_t> roots(n_roots * 4);
    struct Bufs { uint64_t *leaves = nullptr, *scratch = nullptr, *out = nullptr; };
    std::vector<Bufs> bufs(G);
    int rc = HADES_OK;
    auto cleanup = [&]() {
        for (size_t g = 0; g < G; g++) {
            cudaSetDevice(ctx->devs[g].ordinal);
            cudaFree(bufs[g].leaves); cudaFree(bufs[g].scratch); cudaFree(bufs[g].out);
        }
    };
    for (size_t g = 0; g < G && rc == HADES_OK; g++) {
        auto step = [&]() -> int {
            DeviceState& d = ctx->devs[g];
            CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
            CUDA_TRY(ctx, cudaMalloc(&bufs[g].leaves, per_dev * 32));
            CUDA_TRY(ctx, cudaMalloc(&bufs[g].scratch, (per_dev / 4 + per_dev / 16 + 4) * 32));
            CUDA_TRY(ctx, cudaMalloc(&bufs[g].out, std::max<size_t>(roots_per_dev, 1) * 32));
            CUDA_TRY(ctx, cudaMemcpyAsync(bufs[g].leaves, host_leaves + g * per_dev * 4, per_dev * 32,
                                          cudaMemcpyHostToDevice, d.streams[0]));
            int r = merkle_reduce(ctx, bufs[g].leaves, per_dev, sub_levels, bufs[g].scratch, bufs[g].out, d.streams[0]);
            if (r) return r;
            CUDA_TRY(ctx, cudaMemcpyAsync(roots.data() + g * roots_per_dev * 4, bufs[g].out, roots_per_dev * 32,
                                          cudaMemcpyDeviceToHost, d.streams[0]));
            return HADES_OK;
        };
        rc = step();
    }
    for (size_t g = 0; g < G; g++) {
        cudaSetDevice(ctx->devs[g].ordinal);
        cudaError_t e = cudaStreamSynchronize(ctx->devs[g].streams[0]);
        if (e != cudaSuccess && rc == HADES_OK) rc = fail(ctx, HADES_ERR_CUDA, "merkle subtree pass failed: %s", cudaGetErrorString(e));
    }
    if (rc == HADES_OK && n_roots > 1) {
        // top of the tree on the first device (n_roots <= 2*G nodes)
        auto step = [&]() -> int {
            DeviceState& d = ctx->devs[0];
            int top_levels = log4_exact(n_roots);
            if (top_levels < 0) return fail(ctx, HADES_ERR_NOT_POWER_OF_4, "internal: %zu subtree roots", n_roots);
            CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
            CUDA_TRY(ctx, cudaMemcpyAsync(bufs[0].leaves, roots.data(), n_roots * 32, cudaMemcpyHostToDevice, d.streams[0]));
            int r = merkle_reduce(ctx, bufs[0].leaves, n_roots, top_levels, bufs[0].scratch, bufs[0].out, d.streams[0]);
            if (r) return r;
            CUDA_TRY(ctx, cudaMemcpyAsync(roots.data(), bufs[0].out, 32, cudaMemcpyDeviceToHost, d.streams[0]));
            CUDA_TRY(ctx, cudaStreamSynchronize(d.streams[0]));
            return HADES_OK;
        };
        rc = step();
    }
    cleanup();
    if (rc == HADES_OK) memcpy(root, roots.data(), 32);
    return rc;
}

int hades_sponge_batch_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_elems, const uint64_t* d_offsets,
                           size_t n_msgs, uint64_t* d_out, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "sponge needs a width-5 context");
    if (n_msgs == 0) return HADES_OK;
    if (!d_offsets || !d_out) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (((uintptr_t)d_elems | (uintptr_t)d_out) & 15) return fail(ctx, HADES_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    if (n_msgs > 0x7fffffffULL) return fail(ctx, HADES_ERR_INVALID_ARG, "too many messages for one call");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    // Length bucketing: sort message indices by permutation count so that the 32 messages of a warp
    // need the same number of perms (a strictly sequential chain per message).
    struct AsyncBufs {  // stream-ordered temporaries, released on every exit path
        cudaStream_t st;
        void* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        explicit AsyncBufs(cudaStream_t s) : st(s) {}
        ~AsyncBufs() { for (void* q : p) if (q) cudaFreeAsync(q, st); }
    } bufs(st);
    size_t tmp_bytes = 0;
    const int n = (int)n_msgs;
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 32, st));
    for (int i = 0; i < 4; i++) CUDA_TRY(ctx, cudaMallocAsync(&bufs.p[i], n_msgs * 4, st));
    CUDA_TRY(ctx, cudaMallocAsync(&bufs.p[4], tmp_bytes, st));
    uint32_t *keys = (uint32_t*)bufs.p[0], *keys_out = (uint32_t*)bufs.p[1], *idx = (uint32_t*)bufs.p[2],
             *order = (uint32_t*)bufs.p[3];
    sponge_keys_kernel<<<(unsigned)std::min<size_t>((n_msgs + 255) / 256, 148 * 16), 256, 0, st>>>(d_offsets, keys, idx, n_msgs);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(bufs.p[4], tmp_bytes, keys, keys_out, idx, order, n, 0, 32, st));
    ctx->launches++;
    CUDA_TRY(ctx, ctx->ops()->launch_sponge(ctx->variant, d_elems, d_offsets, order, d_out, n_msgs, st));
    return HADES_OK;
}

int hades_sponge_batch(hades_ctx* ctx, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs, uint64_t* out) {
    if (!ctx) return fail(ctx, HADES_ERR_INVALID_ARG, "null context");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "sponge needs a width-5 context");
    if (n_msgs == 0) return HADES_OK;
    if (!offsets || !out) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    for (size_t m = 0; m < n_msgs; m++)
        if (offsets[m + 1] < offsets[m]) return fail(ctx, HADES_ERR_INVALID_ARG, "offsets must be non-decreasing");
    if (offsets[n_msgs] > offsets[0] && !elems) return fail(ctx, HADES_ERR_INVALID_ARG, "null elems pointer");
    // contiguous message ranges per device, balanced by permutation count (floor(len/4) + 1 each)
    const size_t G = std::min<size_t>(ctx->devs.size(), std::max<size_t>(1, n_msgs / 4096));
    std::vector<size_t> bound(G + 1, n_msgs);
    bound[0] = 0;
    if (G > 1) {
        uint64_t total = 0;
        for (size_t m = 0; m < n_msgs; m++) total += (offsets[m + 1] - offsets[m]) / 4 + 1;
        uint64_t acc = 0;
        size_t g = 1;
        for (size_t m = 0; m < n_msgs && g < G; m++) {
            acc += (offsets[m + 1] - offsets[m]) / 4 + 1;
            while (g < G && acc >= total * g / G) bound[g++] = m + 1;
        }
    }
    struct Bufs { uint64_t *elems = nullptr, *offsets = nullptr, *out = nullptr; };
    std::vector<Bufs> bufs(G);
    int rc = HADES_OK;
    for (size_t g = 0; g < G && rc == HADES_OK; g++) {
        auto step = [&]() -> int {
            DeviceState& d = ctx->devs[g];
            const size_t m0 = bound[g], cnt = bound[g + 1] - bound[g];
            if (!cnt) return HADES_OK;
            const uint64_t e0 = offsets[m0], ne = offsets[m0 + cnt] - e0;
            CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
            CUDA_TRY(ctx, cudaMalloc(&bufs[g].elems, std::max<uint64_t>(ne, 1) * 32));
            CUDA_TRY(ctx, cudaMalloc(&bufs[g].offsets, (cnt + 1) * 8));
            CUDA_TRY(ctx, cudaMalloc(&bufs[g].out, cnt * 32));
            if (ne) CUDA_TRY(ctx, cudaMemcpyAsync(bufs[g].elems, elems + e0 * 4, ne * 32, cudaMemcpyHostToDevice, d.streams[0]));
            CUDA_TRY(ctx, cudaMemcpyAsync(bufs[g].offsets, offsets + m0, (cnt + 1) * 8, cudaMemcpyHostToDevice, d.streams[0]));
            // offsets stay absolute: bias the element base pointer by the range's first offset
            int r = hades_sponge_batch_dev(ctx, (int)g, bufs[g].elems - e0 * 4, bufs[g].offsets, cnt, bufs[g].out, d.streams[0]);
            if (r) return r;
            CUDA_TRY(ctx, cudaMemcpyAsync(out + m0 * 4, bufs[g].out, cnt * 32, cudaMemcpyDeviceToHost, d.streams[0]));
            return HADES_OK;
        };
        rc = step();
    }
    for (size_t g = 0; g < G; g++) {
        cudaSetDevice(ctx->devs[g].ordinal);
        cudaError_t e = cudaStreamSynchronize(ctx->devs[g].streams[0]);
        if (e != cudaSuccess && rc == HADES_OK) rc = fail(ctx, HADES_ERR_CUDA, "sponge pass failed: %s", cudaGetErrorString(e));
        cudaFree(bufs[g].elems); cudaFree(bufs[g].offsets); cudaFree(bufs[g].out);
    }
    return rc;
}

int hades_host_register(hades_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    CUDA_TRY(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return HADES_OK;
}
int hades_host_unregister(hades_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    CUDA_TRY(ctx, cudaHostUnregister(ptr));
    return HADES_OK;
}

int hades_gen_elems_dev(hades_ctx* ctx, int dev_index, uint64_t* d_out, uint64_t first_elem, size_t n_elems,
                        uint64_t seed, void* stream) {
    if (!valid_dev(ctx, dev_index) || (!d_out && n_elems)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad argument");
    if (!n_elems) return HADES_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    size_t blocks = std::min<size_t>((n_elems * 4 + 255) / 256, 148 * 16);
    gen_elems_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_out, first_elem, n_elems, seed);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return HADES_OK;
}

int hades_digest_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_limbs, uint64_t first_limb, size_t n_limbs,
                     uint64_t* d_digest, void* stream) {
    if (!valid_dev(ctx, dev_index) || !d_digest || (!d_limbs && n_limbs)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad argument");
    if (!n_limbs) return HADES_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    size_t blocks = std::min<size_t>((n_limbs + 255) / 256, 148 * 16);
    digest_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_limbs, first_limb, n_limbs,
                                                                    reinterpret_cast<unsigned long long*>(d_digest));
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return HADES_OK;
}

int hades_imad_peak(hades_ctx* ctx, int dev_index, int variant, double* products_per_s) {
    if (!valid_dev(ctx, dev_index) || !products_per_s || variant < 0 || variant > 4)
        return fail(ctx, HADES_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    const int threads = 256, blocks = 148 * 8;
    uint32_t *d_in = nullptr, *d_out = nullptr;
    std::vector<uint32_t> h(64 * 32);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)splitmix64(i + 1) | 1u;
    CUDA_TRY(ctx, cudaMalloc(&d_in, h.size() * 4));
    CUDA_TRY(ctx, cudaMalloc(&d_out, (size_t)threads * blocks * 4));
    CUDA_TRY(ctx, cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    auto launch = [&](int iters) {
        switch (variant) {
            case 0: imad_peak_kernel<0><<<blocks, threads>>>(d_in, d_out, iters); break;
            case 1: imad_peak_kernel<1><<<blocks, threads>>>(d_in, d_out, iters); break;
            case 2: imad_peak_kernel<2><<<blocks, threads>>>(d_in, d_out, iters); break;
            case 4: imad_peak_kernel<4><<<blocks, threads>>>(d_in, d_out, iters); break;
            default: imad_peak_kernel<3><<<blocks, threads>>>(d_in, d_out, iters); break;
        }
        ctx->launches++;
    };
    launch(64);  // warm-up
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    const int iters = 1024;
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CUDA_TRY(ctx, cudaEventRecord(e0));
        launch(iters);
        CUDA_TRY(ctx, cudaEventRecord(e1));
        CUDA_TRY(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double prods = (double)threads * blocks * (double)iters * (variant == 0 ? kPeakProductsPerIterV0 : variant == 4 ? kPeakProductsPerIterV4 : kPeakProductsPerIterV123);
        best = std::max(best, prods / (ms * 1e-3));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_in); cudaFree(d_out);
    *products_per_s = best;
    return HADES_OK;
}

int hades_kernel_info(hades_ctx* ctx, const char* kernel, int* regs_per_thread, int* local_bytes,
                      int* max_threads_per_block) {
    if (!ctx || !kernel) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    cudaFuncAttributes a;
    cudaError_t e = ctx->generic() ? (strcmp(kernel, "perm") ? cudaErrorInvalidValue : generic_func_attributes(&a))
                                   : ctx->ops()->func_attributes(kernel, ctx->variant, &a);
    if (e == cudaErrorInvalidValue) return fail(ctx, HADES_ERR_INVALID_ARG, "unknown kernel '%s' for width %u", kernel, ctx->width);
    CUDA_TRY(ctx, e);
    if (regs_per_thread) *regs_per_thread = a.numRegs;
    if (local_bytes) *local_bytes = (int)a.localSizeBytes;
    if (max_threads_per_block) *max_threads_per_block = a.maxThreadsPerBlock;
    return HADES_OK;
}

int hades_set_variant(hades_ctx* ctx, int algo, int regs) {
    if (!ctx || algo < 0 || algo > 2 || regs < 0 || regs > 10) return fail(ctx, HADES_ERR_INVALID_ARG, "variant out of range");
    if (ctx->generic()) return fail(ctx, HADES_ERR_INVALID_ARG, "width %u runs the generic kernel, which has no variants", ctx->width);
    if (algo == 2 && !ctx->has_ccf) return fail(ctx, HADES_ERR_CONSTANTS, "the canonical-form schedule could not be derived for these constants");
    ctx->variant.algo = algo;
    ctx->variant.regs = regs;
    return HADES_OK;
}

int hades_set_coop_threshold(hades_ctx* ctx, size_t max_states) {
    if (!ctx) return fail(ctx, HADES_ERR_INVALID_ARG, "null context");
    if (ctx->width != 5 || !ctx->has_ccf)
        return fail(ctx, HADES_ERR_INVALID_ARG, "the cooperative kernels exist for width 5 with the canonical-form tables only");
    ctx->variant.coop_max = (int)std::min<size_t>(max_states, (size_t)1 << 24);
    return HADES_OK;
}

}  // extern "C"
