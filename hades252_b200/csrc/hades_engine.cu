// libhades_b200.so -- host side of the batched Hades252 engine: context, constant upload,
// multi-device sharding, chunked H2D/compute/D2H pipeline, Merkle and sponge drivers, C ABI
// (include/hades_cuda.h).  No CPU fallback anywhere: every entry point either runs CUDA kernels
// or fails with a status code.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hades_cuda.h"
#include "fr.cuh"
#include "host_tables.hpp"
#include "util_kernels.cuh"
#include "width_ops.hpp"

using namespace hades;

namespace {

constexpr int kRounds = 67;                      // 8 full + 59 partial (src/lib.rs:20-27)
constexpr int kNumBuf = 4;                       // chunk buffers (and streams) per device
constexpr size_t kChunkBytes = (size_t)96 << 20;  // target bytes per pipeline chunk
// Batches / Merkle levels up to this many states run the cooperative 8-lanes-per-state kernels (coop.cuh): below it
// the one-thread-per-state kernel is latency-bound (one warp per scheduler, ~270 us whatever the size).
// Measured (profiles/r02_latency_small_batches.txt): 111-113 us up to 2368 states (one 16-state block per SM), 168 us up to
// 4736, 314 us at 8192 -- against 270 us for the one-thread kernel at any size up to 16 384.
constexpr int kDefaultCoopMax = 4736;
// ... and up to this many the warp-per-state version (one block of four states per SM): 98.6 us instead of 111 us for a
// lone permutation, slower than the 8-lane kernel beyond one block per SM
constexpr int kDefaultCoopWideMax = 592;
// Host batches of at most this many bytes on a single-device context (a lone `Strategy::perm` is 160 B) skip the copy
// engines: the states are memcpy'd into a mapped page-locked buffer and the kernel reads and writes that buffer over
// PCIe.  Two cudaMemcpyAsync calls of ~10 us each are a quarter of the 99 us kernel.
constexpr size_t kTinyBytes = 16 << 10;

struct DeviceState {
    int ordinal = 0;
    uint64_t* generic_tables = nullptr;  // widths without a tuned kernel: ark ++ mds in global memory
    cudaStream_t streams[kNumBuf] = {};
    uint64_t* chunk[kNumBuf] = {};       // device chunk buffers of the host pipeline (grow-only)
    size_t chunk_bytes = 0;
    uint64_t* bounce[kNumBuf] = {};      // pinned staging buffers for PAGEABLE caller memory (grow-only)
    size_t bounce_bytes = 0;
    cudaEvent_t done[kNumBuf] = {};      // chunk b's D2H has landed in bounce[b]
    uint64_t* work = nullptr;            // persistent scratch of the Merkle / sponge host paths (grow-only)
    size_t work_bytes = 0;
    uint64_t* tiny = nullptr;            // mapped page-locked buffer of the tiny-batch host path (kTinyBytes) ...
    uint64_t* tiny_dev = nullptr;        // ... and its device alias: the kernel works on it in place, no copies
    ncclComm_t comm = nullptr;           // rank of this device in the context's communicator (n_dev > 1)
};

// ---- NCCL, loaded on demand (single-device contexts never touch it) ------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool ok() const { return handle != nullptr; }
};
NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the soname every NCCL 2.x ships; a process that already loaded one (e.g. torch's bundled copy) gets that one
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
        bool all = true;
        auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) { all = false; api.error = std::string("missing symbol ") + name; } return p; };
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (all) api.handle = h;
    });
    return api;
}

// ---- host copy pool: slices of user<->pinned copies for the pageable pipeline ----------------------------------
class CopyPool {
  public:
    static CopyPool& get() { static CopyPool p; return p; }
    // dst <- src, split over the workers; returns when every byte has been copied
    void copy(void* dst, const void* src, size_t bytes) {
        const size_t kSlice = (size_t)1 << 19;
        const size_t n = std::max<size_t>(1, std::min<size_t>(workers_.size(), (bytes + kSlice - 1) / kSlice));
        if (n == 1 || workers_.empty()) { memcpy(dst, src, bytes); return; }
        struct Job { std::atomic<size_t> left; std::mutex m; std::condition_variable cv; } job;
        job.left = n;
        const size_t per = ((bytes + n - 1) / n + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (size_t i = 0; i < n; i++) {
                const size_t off = std::min(bytes, i * per), len = std::min(per, bytes - off);
                q_.push_back([=, &job] {
                    if (len) memcpy((char*)dst + off, (const char*)src + off, len);
                    if (job.left.fetch_sub(1) == 1) { std::lock_guard<std::mutex> l2(job.m); job.cv.notify_all(); }
                });
            }
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(job.m);
        job.cv.wait(lk, [&] { return job.left.load() == 0; });
    }
  private:
    CopyPool() {
        // the box's hardware threads are shared by the ranks of a torchrun job (LOCAL_WORLD_SIZE processes)
        unsigned hw = std::thread::hardware_concurrency();
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) hw = std::max(2u, hw / (unsigned)std::max(1, atoi(e)));
        if (const char* e = getenv("HADES_COPY_THREADS")) hw = (unsigned)std::max(0, atoi(e));
        const unsigned n = std::min(16u, hw);
        for (unsigned i = 0; i < n; i++) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
};

// every entry point leaves the caller's current device as it found it
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

thread_local std::string g_init_error;

}  // namespace

struct hades_ctx {
    uint32_t width = 0;
    const WidthOps* ops2[3] = {nullptr, nullptr, nullptr};  // [algo]
    const WidthOps* ops() const { return ops2[variant.algo]; }
    bool generic() const { return ops2[0] == nullptr; }  // no tuned kernel for this width
    Variant variant = {1, 0};  // optimised schedule, <=128 registers
    bool has_sparse = false;   // sparse partial-round tables derived (algo 1 available)
    bool has_ccf = false;      // canonical-form tables derived (algo 2 available)
    std::vector<DeviceState> devs;
    mutable std::string err;
    std::atomic<uint64_t> launches{0};
    bool use_nccl = false;     // communicator over all devices of the context (n_dev > 1, distinct ordinals)
    std::string collective = "none (single device)";
    bool probe = false;        // hades_copy_probe: run the host pipeline without the kernel
    bool force_bounce = false, force_direct = false;  // hades_set_host_path (tests / A-B)
    std::string last_host_path = "none";
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
    static std::mutex err_mutex;  // device pipelines of one call may fail concurrently
    std::lock_guard<std::mutex> lk(err_mutex);
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
    if (!opt.empty()) CUDA_TRY(ctx, ctx->ops2[1]->upload(opt.data()));
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

// grow-only device / pinned buffers owned by the context (the current device must be d.ordinal)
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
int ensure_bounce(hades_ctx* ctx, DeviceState& d, size_t bytes) {
    if (d.bounce_bytes >= bytes) return HADES_OK;
    for (int b = 0; b < kNumBuf; b++) {
        if (d.bounce[b]) CUDA_TRY(ctx, cudaFreeHost(d.bounce[b]));
        d.bounce[b] = nullptr;
    }
    d.bounce_bytes = 0;
    for (int b = 0; b < kNumBuf; b++) CUDA_TRY(ctx, cudaHostAlloc(&d.bounce[b], bytes, cudaHostAllocPortable));
    d.bounce_bytes = bytes;
    return HADES_OK;
}
int ensure_tiny(hades_ctx* ctx, DeviceState& d) {
    if (d.tiny) return HADES_OK;
    void* h = nullptr;
    CUDA_TRY(ctx, cudaHostAlloc(&h, kTinyBytes, cudaHostAllocMapped | cudaHostAllocPortable));
    void* dp = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dp, h, 0);
    if (e != cudaSuccess) {
        cudaFreeHost(h);
        CUDA_TRY(ctx, e);
    }
    d.tiny = static_cast<uint64_t*>(h);
    d.tiny_dev = static_cast<uint64_t*>(dp);
    return HADES_OK;
}
int ensure_work(hades_ctx* ctx, DeviceState& d, size_t bytes) {
    if (d.work_bytes >= bytes) return HADES_OK;
    if (d.work) CUDA_TRY(ctx, cudaFree(d.work));
    d.work = nullptr;
    d.work_bytes = 0;
    bytes = (bytes + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
    CUDA_TRY(ctx, cudaMalloc(&d.work, bytes));
    d.work_bytes = bytes;
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

#define NCCL_TRY(ctx, expr)                                                                               \
    do {                                                                                                  \
        ncclResult_t r_ = (expr);                                                                         \
        if (r_ != ncclSuccess)                                                                            \
            return fail(ctx, HADES_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(r_), \
                        __FILE__, __LINE__);                                                              \
    } while (0)

// One communicator rank per device of the context (single process: ncclCommInitAll).  Without NCCL (library not
// found) or with a device listed twice (NCCL needs distinct devices) the subtree roots travel by direct peer copies.
void init_collective(hades_ctx* ctx) {
    const int G = (int)ctx->devs.size();
    if (G < 2) return;
    std::vector<int> ords;
    for (auto& d : ctx->devs) ords.push_back(d.ordinal);
    std::vector<int> sorted = ords;
    std::sort(sorted.begin(), sorted.end());
    if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) {
        ctx->collective = "peer copies (cudaMemcpyPeerAsync): a device is listed twice, NCCL needs distinct devices";
        return;
    }
    NcclApi& api = nccl_api();
    if (!api.ok()) {
        ctx->collective = "peer copies (cudaMemcpyPeerAsync): " + api.error;
        return;
    }
    std::vector<ncclComm_t> comms(G, nullptr);
    ncclResult_t r = api.CommInitAll(comms.data(), G, ords.data());
    if (r != ncclSuccess) {
        ctx->collective = std::string("peer copies (cudaMemcpyPeerAsync): ncclCommInitAll failed: ") + api.GetErrorString(r);
        return;
    }
    for (int g = 0; g < G; g++) ctx->devs[g].comm = comms[g];
    int ver = 0;
    api.GetVersion(&ver);
    char buf[96];
    snprintf(buf, sizeof buf, "ncclAllGather (NCCL %d.%d.%d, ncclCommInitAll over %d devices)", ver / 10000, (ver / 100) % 100, ver % 100, G);
    ctx->collective = buf;
    ctx->use_nccl = true;
}

// ---- host pipeline of hades_perm_batch on ONE device: states [lo, hi) of the caller's buffer -------------------
// bounce: the caller's memory is PAGEABLE (a Rust `&mut [[BlsScalar; WIDTH]]`, a numpy array).  cudaMemcpyAsync from
// pageable memory is synchronous and staged by the driver on the calling thread, so the chunks are staged here
// instead: the copy pool fills pinned buffer b from the caller's memory, the stream does H2D -> kernel -> D2H in
// place, and a drainer thread copies the results back while the next chunks are in flight.
int pipeline_bounce(hades_ctx* ctx, DeviceState& d, uint64_t* host_states, size_t lo, size_t hi, size_t chunk_states) {
    const size_t state_bytes = (size_t)ctx->width * 32;
    const size_t n_chunks = (hi - lo + chunk_states - 1) / chunk_states;
    std::mutex m;
    std::condition_variable cv;
    size_t issued = 0, drained = 0;  // chunks handed to the stream / copied back to the caller
    bool failed = false;
    cudaError_t drain_err = cudaSuccess;
    std::thread drainer([&] {
        cudaSetDevice(d.ordinal);
        for (size_t c = 0; c < n_chunks; c++) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return issued > c || failed; });
                if (issued <= c) return;
            }
            const int b = (int)(c % kNumBuf);
            cudaError_t e = cudaEventSynchronize(d.done[b]);
            const size_t at = lo + c * chunk_states, cnt = std::min(chunk_states, hi - at);
            if (e == cudaSuccess) CopyPool::get().copy(host_states + at * ctx->width * 4, d.bounce[b], cnt * state_bytes);
            std::lock_guard<std::mutex> lk(m);
            if (e != cudaSuccess) { drain_err = e; failed = true; }
            drained = c + 1;
            cv.notify_all();
        }
    });
    int rc = HADES_OK;
    auto body = [&]() -> int {
        for (size_t c = 0; c < n_chunks; c++) {
            const int b = (int)(c % kNumBuf);
            {
                std::unique_lock<std::mutex> lk(m);  // buffer b is free once chunk c - kNumBuf has been drained
                cv.wait(lk, [&] { return c < kNumBuf || drained + kNumBuf > c || failed; });
                if (failed) return HADES_OK;
            }
            const size_t at = lo + c * chunk_states, cnt = std::min(chunk_states, hi - at);
            CopyPool::get().copy(d.bounce[b], host_states + at * ctx->width * 4, cnt * state_bytes);
            CUDA_TRY(ctx, cudaMemcpyAsync(d.chunk[b], d.bounce[b], cnt * state_bytes, cudaMemcpyHostToDevice, d.streams[b]));
            if (!ctx->probe) {
                int r = launch_perm_w(ctx, d.chunk[b], cnt, d.streams[b], &d);
                if (r) return r;
            }
            CUDA_TRY(ctx, cudaMemcpyAsync(d.bounce[b], d.chunk[b], cnt * state_bytes, cudaMemcpyDeviceToHost, d.streams[b]));
            CUDA_TRY(ctx, cudaEventRecord(d.done[b], d.streams[b]));
            std::lock_guard<std::mutex> lk(m);
            issued = c + 1;
            cv.notify_all();
        }
        return HADES_OK;
    };
    rc = body();
    {
        std::lock_guard<std::mutex> lk(m);
        if (rc != HADES_OK) failed = true;
        cv.notify_all();
    }
    drainer.join();
    if (rc == HADES_OK && drain_err != cudaSuccess)
        rc = fail(ctx, HADES_ERR_CUDA, "perm_batch pipeline failed on device %d: %s", d.ordinal, cudaGetErrorString(drain_err));
    return rc;
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
    DeviceGuard guard;
    hades_ctx* ctx = new hades_ctx();
    ctx->width = width;
    if (tuned)
        for (int a = 0; a < 3; a++) ctx->ops2[a] = width == 3 ? width_ops_3(a) : width == 5 ? width_ops_5(a) : width_ops_9(a);
    ctx->variant.algo = 0;
    ctx->variant.regs = 0;
    // dense table = ROUND_CONSTANTS[0..67W) ++ MDS_MATRIX; optimised table derived from it (host_tables.hpp)
    std::vector<uint64_t> dense(ark_limbs, ark_limbs + (size_t)kRounds * width * 4), opt, ccf;
    dense.insert(dense.end(), mds_limbs, mds_limbs + (size_t)width * width * 4);
    int rc = HADES_OK;
    if (tuned) {
        if (dense.size() != ctx->ops2[0]->table_u64) rc = fail(nullptr, HADES_ERR_INVALID_ARG, "internal: dense table size");
        // Schedules, best first: gauged canonical form (needs a controllable MDS block and non-zero pivots), sparse
        // partial rounds (needs invertible MDS sub-matrices), dense (the reference's round structure, always
        // available).  With the reference's assets all three derive; with custom constants where a derivation
        // fails the next one becomes the default -- the dense schedule needs no derived table at all.
        if (rc == HADES_OK) {
            ctx->has_sparse = hades_host::derive_tables((int)width, ark_limbs, mds_limbs, opt) && opt.size() == ctx->ops2[1]->table_u64;
            if (!ctx->has_sparse) opt.clear();
            ctx->has_ccf = hades_host::derive_tables_ccf((int)width, ark_limbs, mds_limbs, ccf) && ccf.size() == ctx->ops2[2]->table_u64;
            if (!ctx->has_ccf) ccf.clear();
            if (ctx->has_sparse || ctx->has_ccf) {
                ctx->variant.algo = ctx->has_ccf ? 2 : 1;
                ctx->variant.regs = width == 9 ? 7 : 6;  // lockstep 128-thread blocks: x7 (W=3), x5 (W=5), x3 (W=9) per SM
                if (ctx->has_ccf && width == 5) {
                    ctx->variant.coop_max = kDefaultCoopMax;
                    ctx->variant.coop_wide_max = kDefaultCoopWideMax;
                }
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
            for (int b = 0; b < kNumBuf; b++) {
                CUDA_TRY(ctx, cudaStreamCreateWithFlags(&d.streams[b], cudaStreamNonBlocking));
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&d.done[b], cudaEventDisableTiming));
            }
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
    init_collective(ctx);
    *out = ctx;
    return HADES_OK;
}

void hades_destroy(hades_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard;
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.ordinal);
        for (int b = 0; b < kNumBuf; b++)
            if (d.streams[b]) cudaStreamSynchronize(d.streams[b]);
        if (d.comm) nccl_api().CommDestroy(d.comm);
        for (int b = 0; b < kNumBuf; b++) {
            if (d.chunk[b]) cudaFree(d.chunk[b]);
            if (d.bounce[b]) cudaFreeHost(d.bounce[b]);
            if (d.done[b]) cudaEventDestroy(d.done[b]);
            if (d.streams[b]) cudaStreamDestroy(d.streams[b]);
        }
        if (d.work) cudaFree(d.work);
        if (d.tiny) cudaFreeHost(d.tiny);
        if (d.generic_tables) cudaFree(d.generic_tables);
    }
    delete ctx;
}

const char* hades_last_error(const hades_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }
uint32_t hades_width(const hades_ctx* ctx) { return ctx ? ctx->width : 0; }
int hades_device_count(const hades_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
uint64_t hades_launch_count(const hades_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
const char* hades_collective(const hades_ctx* ctx) { return ctx ? ctx->collective.c_str() : ""; }
const char* hades_last_host_path(const hades_ctx* ctx) { return ctx ? ctx->last_host_path.c_str() : ""; }

int hades_perm_batch_dev(hades_ctx* ctx, int dev_index, uint64_t* d_states, size_t n, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    if (n == 0) return HADES_OK;
    if (!d_states || ((uintptr_t)d_states & 15)) return fail(ctx, HADES_ERR_INVALID_ARG, "d_states must be non-null and 16-byte aligned");
    DeviceGuard guard;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    return launch_perm_w(ctx, d_states, n, (cudaStream_t)stream, &ctx->devs[dev_index]);
}

int hades_perm_batch(hades_ctx* ctx, uint64_t* host_states, size_t n) {
    if (!ctx) return fail(ctx, HADES_ERR_INVALID_ARG, "null context");
    if (n == 0) return HADES_OK;
    if (!host_states) return fail(ctx, HADES_ERR_INVALID_ARG, "null states pointer");
    DeviceGuard guard;
    const size_t state_bytes = (size_t)ctx->width * 32;
    const size_t G = ctx->devs.size();
    if (G == 1 && n * state_bytes <= kTinyBytes && !ctx->probe && !ctx->force_bounce && !ctx->force_direct) {
        // tiny batch (a lone `Strategy::perm`, src/strategies.rs:140): in place on a mapped page-locked buffer
        DeviceState& d = ctx->devs[0];
        CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
        int r = ensure_tiny(ctx, d);
        if (r) return r;
        memcpy(d.tiny, host_states, n * state_bytes);
        r = launch_perm_w(ctx, d.tiny_dev, n, d.streams[0], &d);
        if (r) return r;
        CUDA_TRY(ctx, cudaStreamSynchronize(d.streams[0]));
        memcpy(host_states, d.tiny, n * state_bytes);
        ctx->last_host_path = "tiny batch: the kernel works in place on a mapped page-locked buffer (no copy engines)";
        return HADES_OK;
    }
    size_t chunk_states = std::max<size_t>(kPermThreads, kChunkBytes / state_bytes / kPermThreads * kPermThreads);
    // Medium batches (at least kMinSplitBytes per chunk) are split into kNumBuf chunks so that H2D, kernel and D2H --
    // and, for pageable memory, the staging copies -- overlap.  Smaller batches go as ONE chunk: a launch is one
    // latency-bound wave whatever its size up to ~16k states, so splitting would only add launches.
    // (HADES_MIN_SPLIT_BYTES overrides the threshold: measurement knob.)
    static const size_t kMinSplitBytes = [] {
        const char* e = getenv("HADES_MIN_SPLIT_BYTES");
        return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)1 << 20;
    }();
    const size_t per_dev = (n + G - 1) / G;
    if (per_dev < chunk_states * kNumBuf) {
        if (per_dev * state_bytes >= kMinSplitBytes * kNumBuf)
            chunk_states = std::max<size_t>(kPermThreads, (per_dev / kNumBuf + kPermThreads) / kPermThreads * kPermThreads);
        else
            chunk_states = std::max<size_t>(kPermThreads, (per_dev + kPermThreads - 1) / kPermThreads * kPermThreads);
    }
    // Page-locked caller memory (cudaHostAlloc, hades_host_register) is copied from directly; PAGEABLE memory of more
    // than one chunk goes through the context's pinned staging buffers (pipeline_bounce); a small pageable batch is
    // one synchronous staged copy either way.
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, host_states) == cudaSuccess &&
                  (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    bool bounce = !pinned && per_dev > chunk_states;
    if (ctx->force_bounce) bounce = true;
    if (ctx->force_direct) bounce = false;
    ctx->last_host_path = bounce ? "pageable: staged through pinned bounce buffers by the copy pool"
                          : pinned ? "page-locked: direct asynchronous copies" : "pageable: single synchronous staged copy";
    std::vector<size_t> lo(G), hi(G);
    for (size_t g = 0; g < G; g++) {
        lo[g] = n * g / G;
        hi[g] = n * (g + 1) / G;
        if (hi[g] > lo[g]) {
            CUDA_TRY(ctx, cudaSetDevice(ctx->devs[g].ordinal));
            const size_t bytes = std::min(chunk_states, hi[g] - lo[g]) * state_bytes;
            int r = ensure_chunks(ctx, ctx->devs[g], bytes);
            if (r == HADES_OK && bounce) r = ensure_bounce(ctx, ctx->devs[g], bytes);
            if (r) return r;
        }
    }
    int rc = HADES_OK;
    if (bounce) {  // one pipeline thread per device (each blocks on its own staging copies)
        std::vector<int> rcs(G, HADES_OK);
        std::vector<std::thread> th;
        for (size_t g = 1; g < G; g++)
            if (hi[g] > lo[g])
                th.emplace_back([&, g] {
                    cudaSetDevice(ctx->devs[g].ordinal);
                    rcs[g] = pipeline_bounce(ctx, ctx->devs[g], host_states, lo[g], hi[g], chunk_states);
                });
        if (hi[0] > lo[0]) {
            cudaSetDevice(ctx->devs[0].ordinal);
            rcs[0] = pipeline_bounce(ctx, ctx->devs[0], host_states, lo[0], hi[0], chunk_states);
        }
        for (auto& t : th) t.join();
        for (int r : rcs) if (r != HADES_OK && rc == HADES_OK) rc = r;
    } else {  // asynchronous copies: issue round-robin over the devices from this thread
        size_t max_chunks = 0;
        for (size_t g = 0; g < G; g++) max_chunks = std::max(max_chunks, (hi[g] - lo[g] + chunk_states - 1) / chunk_states);
        for (size_t c = 0; c < max_chunks && rc == HADES_OK; c++)
            for (size_t g = 0; g < G && rc == HADES_OK; g++) {
                const size_t at = lo[g] + c * chunk_states;
                if (at >= hi[g]) continue;
                auto step = [&]() -> int {
                    DeviceState& d = ctx->devs[g];
                    CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
                    // chunk c reuses buffer c % kNumBuf of its own stream: stream order protects the buffer
                    const size_t cnt = std::min(chunk_states, hi[g] - at);
                    const int b = (int)(c % kNumBuf);
                    uint64_t* h = host_states + at * ctx->width * 4;
                    CUDA_TRY(ctx, cudaMemcpyAsync(d.chunk[b], h, cnt * state_bytes, cudaMemcpyHostToDevice, d.streams[b]));
                    if (!ctx->probe) {
                        int r = launch_perm_w(ctx, d.chunk[b], cnt, d.streams[b], &d);
                        if (r) return r;
                    }
                    CUDA_TRY(ctx, cudaMemcpyAsync(h, d.chunk[b], cnt * state_bytes, cudaMemcpyDeviceToHost, d.streams[b]));
                    return HADES_OK;
                };
                rc = step();
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

int hades_copy_probe(hades_ctx* ctx, uint64_t* host_states, size_t n) {
    if (!ctx) return fail(ctx, HADES_ERR_INVALID_ARG, "null context");
    ctx->probe = true;
    int rc = hades_perm_batch(ctx, host_states, n);
    ctx->probe = false;
    return rc;
}

int hades_set_host_path(hades_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return fail(ctx, HADES_ERR_INVALID_ARG, "mode must be 0 (auto), 1 (bounce) or 2 (direct)");
    ctx->force_bounce = mode == 1;
    ctx->force_direct = mode == 2;
    return HADES_OK;
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
    DeviceGuard guard;
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
    DeviceGuard guard;
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
    DeviceGuard guard;
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

int hades_merkle_verify_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_leaves, size_t n_leaves, const uint64_t* d_index,
                            size_t n_open, const uint64_t* d_branch, const uint64_t* d_root, uint32_t* d_ok, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "merkle needs a width-5 context");
    if (n_leaves == 0 || !d_leaves) return fail(ctx, HADES_ERR_INVALID_ARG, "a tree needs at least one leaf");
    if (n_open == 0) return HADES_OK;
    int levels = 0;
    for (size_t m = n_leaves; m > 1; m = (m + 3) / 4) levels++;
    if (!d_index || !d_root || !d_ok || (levels && !d_branch)) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (((uintptr_t)d_leaves | (uintptr_t)d_branch | (uintptr_t)d_root) & 15)
        return fail(ctx, HADES_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    DeviceGuard guard;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    ctx->launches++;
    CUDA_TRY(ctx, ctx->ops()->launch_merkle_verify(ctx->variant, d_leaves, d_index, n_open, n_leaves, levels, d_branch, d_root, d_ok,
                                                   (cudaStream_t)stream));
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
    DeviceGuard guard;
    DeviceState& d = ctx->devs[0];
    const size_t nodes = hades_merkle_tree_nodes(n_leaves);
    CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
    int r = ensure_work(ctx, d, (n_leaves + nodes) * 32);  // persistent scratch: leaves | interior levels
    if (r) return r;
    uint64_t* d_leaves = d.work;
    uint64_t* d_tree = d.work + n_leaves * 4;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_leaves, host_leaves, n_leaves * 32, cudaMemcpyHostToDevice, d.streams[0]));
    r = hades_merkle_tree_dev(ctx, 0, d_leaves, n_leaves, d_tree, d.streams[0]);
    if (r) return r;
    CUDA_TRY(ctx, cudaMemcpyAsync(root, d_tree + (nodes - 1) * 4, 32, cudaMemcpyDeviceToHost, d.streams[0]));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.streams[0]));
    return HADES_OK;
}

// Sharded root: device g reduces leaves [g n/G, (g+1) n/G) to 1-2 subtree roots, the roots are
// all-gathered (ncclAllGather over the context's communicator, in place: 32-64 B per device), and every device
// finishes the top levels redundantly; the root is read back from the first device.  All devices are issued
// before anything is waited for; buffers are the context's persistent scratch.
// host_leaves != nullptr: leaves are uploaded from the host; else d_leaves[g] = device g's resident leaf range.
static int merkle_root_sharded(hades_ctx* ctx, const uint64_t* host_leaves, const uint64_t* const* d_leaves, size_t n_leaves,
                               uint64_t root[4]) {
    int depth = log4_exact(n_leaves);
    if (depth < 0) return fail(ctx, HADES_ERR_NOT_POWER_OF_4, "number of leaves (%zu) must be a power of 4", n_leaves);
    DeviceGuard guard;
    const size_t Gall = ctx->devs.size();
    // shard over ALL devices of the context when their number is a power of two and every device gets at least 1024
    // leaves (the communicator spans all of them); otherwise the first device does the whole tree
    size_t G = 1;
    if (Gall > 1 && (Gall & (Gall - 1)) == 0 && n_leaves / Gall >= 1024) G = Gall;
    if (!host_leaves && G != Gall) return fail(ctx, HADES_ERR_INVALID_ARG, "resident leaves need a power-of-two number of devices and >= 1024 leaves each");
    const size_t per_dev = n_leaves / G;                 // = 4^a or 2*4^a
    int sub_levels = 0;                                  // levels each device can reduce on its own range
    while ((per_dev >> (2 * (sub_levels + 1))) >= 1 && ((per_dev >> (2 * (sub_levels + 1))) << (2 * (sub_levels + 1))) == per_dev)
        sub_levels++;
    const size_t roots_per_dev = per_dev >> (2 * sub_levels);  // 1 or 2
    const size_t n_roots = roots_per_dev * G;
    const int top_levels = log4_exact(n_roots);
    if (top_levels < 0) return fail(ctx, HADES_ERR_NOT_POWER_OF_4, "internal: %zu subtree roots", n_roots);
    // scratch layout per device (32-byte elements): [leaves] | ping-pong levels | gathered roots | result
    const size_t leaf_elems = host_leaves ? per_dev : 0;
    const size_t scratch_elems = per_dev / 4 + per_dev / 16 + 8;
    const size_t total_elems = leaf_elems + scratch_elems + n_roots + 8 + 1;
    int rc = HADES_OK;
    for (size_t g = 0; g < G && rc == HADES_OK; g++) {
        auto step = [&]() -> int {
            DeviceState& d = ctx->devs[g];
            CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
            return ensure_work(ctx, d, total_elems * 32);
        };
        rc = step();
    }
    // Leaves coming from the host: the upload is pipelined with the FIRST level (which holds 3/4 of all permutations):
    // the range of every device is cut into chunks, chunk c is uploaded on stream c % kNumBuf -- from page-locked memory
    // directly, from pageable memory through the pinned bounce buffers filled by the copy pool -- and hashed to its
    // level-1 nodes on the same stream while the next chunks are still in flight.  Chunk-major over the devices.
    int first_level = 0;  // levels < first_level have been issued by the upload pipeline
    if (rc == HADES_OK && host_leaves) {
        constexpr size_t kLeafChunk = (size_t)1 << 19;  // leaves per chunk (16 MB): a multiple of 4
        const bool chunked = sub_levels >= 1 && per_dev >= 2 * kLeafChunk;
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, host_leaves) == cudaSuccess &&
                            (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        cudaGetLastError();
        if (!chunked) {
            for (size_t g = 0; g < G && rc == HADES_OK; g++) {
                auto step = [&]() -> int {
                    DeviceState& d = ctx->devs[g];
                    CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
                    CUDA_TRY(ctx, cudaMemcpyAsync(d.work, host_leaves + g * per_dev * 4, per_dev * 32, cudaMemcpyHostToDevice, d.streams[0]));
                    return HADES_OK;
                };
                rc = step();
            }
        } else {
            const size_t n_chunks = per_dev / kLeafChunk;  // per_dev is 4^a or 2 * 4^a >= 2^20: an exact multiple
            std::vector<char> bounce_used(G * kNumBuf, 0);
            for (size_t g = 0; g < G && rc == HADES_OK && !pinned; g++) {
                cudaSetDevice(ctx->devs[g].ordinal);
                rc = ensure_bounce(ctx, ctx->devs[g], kLeafChunk * 32);
            }
            for (size_t c = 0; c < n_chunks && rc == HADES_OK; c++)
                for (size_t g = 0; g < G && rc == HADES_OK; g++) {
                    auto step = [&]() -> int {
                        DeviceState& d = ctx->devs[g];
                        const int b = (int)(c % kNumBuf);
                        CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
                        uint64_t* leaves = d.work + c * kLeafChunk * 4;
                        const uint64_t* src = host_leaves + (g * per_dev + c * kLeafChunk) * 4;
                        if (pinned) {
                            CUDA_TRY(ctx, cudaMemcpyAsync(leaves, src, kLeafChunk * 32, cudaMemcpyHostToDevice, d.streams[b]));
                        } else {
                            if (bounce_used[g * kNumBuf + b]) CUDA_TRY(ctx, cudaEventSynchronize(d.done[b]));  // its H2D has drained
                            CopyPool::get().copy(d.bounce[b], src, kLeafChunk * 32);
                            CUDA_TRY(ctx, cudaMemcpyAsync(leaves, d.bounce[b], kLeafChunk * 32, cudaMemcpyHostToDevice, d.streams[b]));
                            CUDA_TRY(ctx, cudaEventRecord(d.done[b], d.streams[b]));
                            bounce_used[g * kNumBuf + b] = 1;
                        }
                        // level 1 of this chunk: scratch buffer A (or the device's root slot when the range has one level)
                        uint64_t* scratch = d.work + leaf_elems * 4;
                        uint64_t* gathered = scratch + scratch_elems * 4;
                        uint64_t* out = (sub_levels == 1 ? gathered + g * roots_per_dev * 4 : scratch) + c * (kLeafChunk / 4) * 4;
                        ctx->launches++;
                        CUDA_TRY(ctx, ctx->ops()->launch_merkle_level(ctx->variant, leaves, out, kLeafChunk / 4, kLeafChunk, d.streams[b]));
                        return HADES_OK;
                    };
                    rc = step();
                }
            // the remaining levels run on stream 0: it waits for the other upload streams
            for (size_t g = 0; g < G && rc == HADES_OK; g++) {
                auto step = [&]() -> int {
                    DeviceState& d = ctx->devs[g];
                    CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
                    for (int b = 1; b < kNumBuf; b++) {
                        CUDA_TRY(ctx, cudaEventRecord(d.done[b], d.streams[b]));
                        CUDA_TRY(ctx, cudaStreamWaitEvent(d.streams[0], d.done[b], 0));
                    }
                    return HADES_OK;
                };
                rc = step();
            }
            first_level = 1;
        }
    }
    // Levels are issued LEVEL-MAJOR (for every level: all devices), so that every device has its first, longest
    // kernel queued after G launches instead of after (g * levels) launches of the devices before it.
    for (int l = first_level; l < std::max(sub_levels, 1) && rc == HADES_OK; l++) {
        for (size_t g = 0; g < G && rc == HADES_OK; g++) {
            auto step = [&]() -> int {
                DeviceState& d = ctx->devs[g];
                CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
                uint64_t* scratch = d.work + leaf_elems * 4;
                uint64_t* gathered = scratch + scratch_elems * 4;
                const uint64_t* leaves = host_leaves ? d.work : d_leaves[g];
                // own subtree roots land in their slot of the gather buffer (in-place all-gather)
                uint64_t* own = gathered + g * roots_per_dev * 4;
                if (sub_levels == 0) {
                    CUDA_TRY(ctx, cudaMemcpyAsync(own, leaves, roots_per_dev * 32, cudaMemcpyDeviceToDevice, d.streams[0]));
                    return HADES_OK;
                }
                uint64_t* bufA = scratch;
                uint64_t* bufB = scratch + (per_dev / 4) * 4;
                const size_t n_in = per_dev >> (2 * l), n_out = n_in / 4;
                const uint64_t* in = l == 0 ? leaves : (((l - 1) & 1) ? bufB : bufA);
                uint64_t* out = (l == sub_levels - 1) ? own : ((l & 1) ? bufB : bufA);
                ctx->launches++;
                CUDA_TRY(ctx, ctx->ops()->launch_merkle_level(ctx->variant, in, out, n_out, n_in, d.streams[0]));
                return HADES_OK;
            };
            rc = step();
        }
    }
    auto gathered_of = [&](size_t g) { return ctx->devs[g].work + (leaf_elems + scratch_elems) * 4; };
    if (rc == HADES_OK && G > 1) {
        auto gather = [&]() -> int {
            if (ctx->use_nccl) {
                NcclApi& api = nccl_api();
                NCCL_TRY(ctx, api.GroupStart());
                for (size_t g = 0; g < G; g++) {
                    uint64_t* buf = gathered_of(g);
                    ncclResult_t r = api.AllGather(buf + g * roots_per_dev * 4, buf, roots_per_dev * 4, ncclUint64, ctx->devs[g].comm,
                                                   ctx->devs[g].streams[0]);
                    if (r != ncclSuccess) {
                        api.GroupEnd();
                        return fail(ctx, HADES_ERR_CUDA, "ncclAllGather failed: %s", api.GetErrorString(r));
                    }
                }
                NCCL_TRY(ctx, api.GroupEnd());
            } else {  // no communicator (a device listed twice, or no NCCL library): direct peer copies
                for (size_t g = 0; g < G; g++) {
                    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[g].ordinal));
                    CUDA_TRY(ctx, cudaEventRecord(ctx->devs[g].done[0], ctx->devs[g].streams[0]));
                }
                for (size_t g = 0; g < G; g++) {
                    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[g].ordinal));
                    for (size_t h = 0; h < G; h++) {
                        if (h == g) continue;
                        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->devs[g].streams[0], ctx->devs[h].done[0], 0));
                        CUDA_TRY(ctx, cudaMemcpyPeerAsync(gathered_of(g) + h * roots_per_dev * 4, ctx->devs[g].ordinal,
                                                          gathered_of(h) + h * roots_per_dev * 4, ctx->devs[h].ordinal,
                                                          roots_per_dev * 32, ctx->devs[g].streams[0]));
                    }
                }
            }
            return HADES_OK;
        };
        rc = gather();
    }
    // top levels on every device (redundantly); result read from the first
    for (size_t g = 0; g < G && rc == HADES_OK; g++) {
        auto step = [&]() -> int {
            DeviceState& d = ctx->devs[g];
            CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
            uint64_t* scratch = d.work + leaf_elems * 4;
            uint64_t* gathered = gathered_of(g);
            uint64_t* result = gathered + (n_roots + 8) * 4;
            if (top_levels == 0) {
                CUDA_TRY(ctx, cudaMemcpyAsync(result, gathered, 32, cudaMemcpyDeviceToDevice, d.streams[0]));
            } else {
                int r = merkle_reduce(ctx, gathered, n_roots, top_levels, scratch, result, d.streams[0]);
                if (r) return r;
            }
            if (g == 0) CUDA_TRY(ctx, cudaMemcpyAsync(root, result, 32, cudaMemcpyDeviceToHost, d.streams[0]));
            return HADES_OK;
        };
        rc = step();
    }
    for (size_t g = 0; g < G; g++) {
        cudaSetDevice(ctx->devs[g].ordinal);
        cudaError_t e = cudaStreamSynchronize(ctx->devs[g].streams[0]);
        if (e != cudaSuccess && rc == HADES_OK) rc = fail(ctx, HADES_ERR_CUDA, "merkle pass failed on device %d: %s", ctx->devs[g].ordinal, cudaGetErrorString(e));
    }
    return rc;
}

int hades_merkle_root(hades_ctx* ctx, const uint64_t* host_leaves, size_t n_leaves, uint64_t root[4]) {
    if (!ctx || !host_leaves || !root) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "merkle needs a width-5 context");
    if (n_leaves == 1) {
        memcpy(root, host_leaves, 32);
        return HADES_OK;
    }
    return merkle_root_sharded(ctx, host_leaves, nullptr, n_leaves, root);
}

int hades_merkle_root_sharded_dev(hades_ctx* ctx, const uint64_t* const* d_leaves, size_t n_leaves, uint64_t root[4]) {
    if (!ctx || !d_leaves || !root) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "merkle needs a width-5 context");
    for (size_t g = 0; g < ctx->devs.size(); g++)
        if (!d_leaves[g] || ((uintptr_t)d_leaves[g] & 15)) return fail(ctx, HADES_ERR_INVALID_ARG, "d_leaves[%zu] must be non-null and 16-byte aligned", g);
    return merkle_root_sharded(ctx, nullptr, d_leaves, n_leaves, root);
}

// capacity word of the sponge: zero, or the caller's domain tag (a canonical field element, Montgomery limbs)
static int make_tag(hades_ctx* ctx, const uint64_t* tag, SpongeTag& out) {
    memset(&out, 0, sizeof out);
    if (!tag) return HADES_OK;
    static const uint64_t kP[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    bool below = false;
    for (int i = 3; i >= 0 && !below; i--) {
        if (tag[i] < kP[i]) below = true;
        else if (tag[i] > kP[i]) break;
    }
    if (!below) return fail(ctx, HADES_ERR_INVALID_ARG, "the domain tag must be a canonical field element (< p)");
    for (int k = 0; k < 8; k++) out.l[k] = (uint32_t)(tag[k / 2] >> (32 * (k & 1)));
    return HADES_OK;
}

static int sponge_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_elems, const uint64_t* d_offsets,
                      size_t n_msgs, uint64_t* d_out, const uint64_t* tag, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    SpongeTag stag;
    if (int r = make_tag(ctx, tag, stag)) return r;
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "sponge needs a width-5 context");
    if (n_msgs == 0) return HADES_OK;
    if (!d_offsets || !d_out) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    if (((uintptr_t)d_elems | (uintptr_t)d_out) & 15) return fail(ctx, HADES_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    if (n_msgs > 0x7fffffffULL) return fail(ctx, HADES_ERR_INVALID_ARG, "too many messages for one call");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    if (ctx->variant.algo == 2 && n_msgs <= (size_t)ctx->variant.coop_max) {
        // few messages: the cooperative kernel (one message per 8 lanes) needs no bucketing -- one launch, no temporaries
        ctx->launches++;
        CUDA_TRY(ctx, ctx->ops()->launch_sponge(ctx->variant, d_elems, d_offsets, nullptr, d_out, n_msgs, stag, st));
        return HADES_OK;
    }
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
    ctx->launches += 2;  // the radix sort above and the sponge kernel
    CUDA_TRY(ctx, ctx->ops()->launch_sponge(ctx->variant, d_elems, d_offsets, order, d_out, n_msgs, stag, st));
    return HADES_OK;
}

int hades_sponge_batch_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_elems, const uint64_t* d_offsets,
                           size_t n_msgs, uint64_t* d_out, void* stream) {
    return sponge_dev(ctx, dev_index, d_elems, d_offsets, n_msgs, d_out, nullptr, stream);
}
int hades_sponge_batch_ds_dev(hades_ctx* ctx, int dev_index, const uint64_t* d_elems, const uint64_t* d_offsets,
                              size_t n_msgs, const uint64_t domain_tag[4], uint64_t* d_out, void* stream) {
    if (!domain_tag) return fail(ctx, HADES_ERR_INVALID_ARG, "null domain tag");
    return sponge_dev(ctx, dev_index, d_elems, d_offsets, n_msgs, d_out, domain_tag, stream);
}

static int sponge_host(hades_ctx* ctx, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs, uint64_t* out,
                       const uint64_t* tag) {
    if (!ctx) return fail(ctx, HADES_ERR_INVALID_ARG, "null context");
    if (ctx->width != 5) return fail(ctx, HADES_ERR_INVALID_ARG, "sponge needs a width-5 context");
    if (n_msgs == 0) return HADES_OK;
    if (!offsets || !out) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    for (size_t m = 0; m < n_msgs; m++)
        if (offsets[m + 1] < offsets[m]) return fail(ctx, HADES_ERR_INVALID_ARG, "offsets must be non-decreasing");
    if (offsets[n_msgs] > offsets[0] && !elems) return fail(ctx, HADES_ERR_INVALID_ARG, "null elems pointer");
    {   // tiny call (a lone hash): the one-launch cooperative kernel works on a mapped page-locked buffer, no copy engines
        const uint64_t e0 = offsets[0], ne = offsets[n_msgs] - e0;
        const size_t elems_u64 = (size_t)std::max<uint64_t>(ne, 1) * 4, out_u64 = n_msgs * 4;
        if (ctx->variant.algo == 2 && n_msgs <= (size_t)ctx->variant.coop_max && ne <= kTinyBytes / 32 &&
            (elems_u64 + out_u64 + n_msgs + 1) * 8 <= kTinyBytes) {
            DeviceGuard guard;
            DeviceState& d = ctx->devs[0];
            CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
            int r = ensure_tiny(ctx, d);
            if (r) return r;
            if (ne) memcpy(d.tiny, elems + e0 * 4, ne * 32);
            memcpy(d.tiny + elems_u64 + out_u64, offsets, (n_msgs + 1) * 8);
            // offsets stay absolute: bias the element base pointer by the first offset
            r = sponge_dev(ctx, 0, d.tiny_dev - e0 * 4, d.tiny_dev + elems_u64 + out_u64, n_msgs, d.tiny_dev + elems_u64, tag,
                           d.streams[0]);
            if (r) return r;
            CUDA_TRY(ctx, cudaStreamSynchronize(d.streams[0]));
            memcpy(out, d.tiny + elems_u64, n_msgs * 32);
            return HADES_OK;
        }
    }
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
    int rc = HADES_OK;
    DeviceGuard guard;
    // Large ranges are PIPELINED: the range of a device is cut into chunks of at most kChunkElems elements / kChunkMsgs
    // messages; chunk c is staged into pinned buffer c % kNumBuf (elements | offsets) by the copy pool, uploaded, bucketed and
    // hashed on stream c % kNumBuf, and its digests come back through the same pinned buffer while the next chunks are in
    // flight.  A range that fits one chunk keeps the single-shot path (lowest latency for small calls).
    constexpr uint64_t kChunkElems = (uint64_t)1 << 21;  // 64 MB of elements
    constexpr size_t kChunkMsgs = (size_t)1 << 19;
    auto single_shot = [&](size_t g, size_t m0, size_t cnt) -> int {
        DeviceState& d = ctx->devs[g];
        if (!cnt) return HADES_OK;
        const uint64_t e0 = offsets[m0], ne = offsets[m0 + cnt] - e0;
        CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
        // persistent scratch: elements | digests | offsets
        const size_t elems_u64 = std::max<uint64_t>(ne, 1) * 4, out_u64 = cnt * 4;
        int r = ensure_work(ctx, d, (elems_u64 + out_u64 + cnt + 1) * 8);
        if (r) return r;
        uint64_t* d_elems = d.work;
        uint64_t* d_out = d_elems + elems_u64;
        uint64_t* d_offsets = d_out + out_u64;
        if (ne) CUDA_TRY(ctx, cudaMemcpyAsync(d_elems, elems + e0 * 4, ne * 32, cudaMemcpyHostToDevice, d.streams[0]));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_offsets, offsets + m0, (cnt + 1) * 8, cudaMemcpyHostToDevice, d.streams[0]));
        // offsets stay absolute: bias the element base pointer by the range's first offset
        r = sponge_dev(ctx, (int)g, d_elems - e0 * 4, d_offsets, cnt, d_out, tag, d.streams[0]);
        if (r) return r;
        CUDA_TRY(ctx, cudaMemcpyAsync(out + m0 * 4, d_out, cnt * 32, cudaMemcpyDeviceToHost, d.streams[0]));
        return HADES_OK;
    };
    auto pipelined = [&](size_t g, const std::vector<size_t>& cut) -> int {  // cut: chunk boundaries (message indices)
        DeviceState& d = ctx->devs[g];
        CUDA_TRY(ctx, cudaSetDevice(d.ordinal));
        // one slot per in-flight chunk, identical layout on the device and in the pinned buffer: elements | offsets | digests
        const size_t elems_u64 = kChunkElems * 4 + 40 * 4, off_u64 = kChunkMsgs + 8, out_u64 = kChunkMsgs * 4;
        const size_t slot_u64 = (elems_u64 + off_u64 + out_u64 + 1) & ~(size_t)1;  // 16-byte aligned slots
        int r = ensure_work(ctx, d, slot_u64 * 8 * kNumBuf);
        if (r == HADES_OK) r = ensure_bounce(ctx, d, slot_u64 * 8);
        if (r) return r;
        const size_t n_chunks = cut.size() - 1;
        auto drain = [&](size_t c) -> int {  // digests of chunk c: pinned buffer -> caller
            const int b = (int)(c % kNumBuf);
            CUDA_TRY(ctx, cudaEventSynchronize(d.done[b]));
            const size_t m0 = cut[c], cnt = cut[c + 1] - cut[c];
            CopyPool::get().copy(out + m0 * 4, d.bounce[b] + elems_u64 + off_u64, cnt * 32);
            return HADES_OK;
        };
        for (size_t c = 0; c < n_chunks; c++) {
            const int b = (int)(c % kNumBuf);
            if (c >= (size_t)kNumBuf) {
                r = drain(c - kNumBuf);
                if (r) return r;
            }
            const size_t m0 = cut[c], cnt = cut[c + 1] - cut[c];
            const uint64_t e0 = offsets[m0], ne = offsets[m0 + cnt] - e0;
            uint64_t* h = d.bounce[b];
            uint64_t* dev = d.work + (size_t)b * slot_u64;
            if (ne) CopyPool::get().copy(h, elems + e0 * 4, ne * 32);
            memcpy(h + elems_u64, offsets + m0, (cnt + 1) * 8);
            if (ne) CUDA_TRY(ctx, cudaMemcpyAsync(dev, h, ne * 32, cudaMemcpyHostToDevice, d.streams[b]));
            CUDA_TRY(ctx, cudaMemcpyAsync(dev + elems_u64, h + elems_u64, (cnt + 1) * 8, cudaMemcpyHostToDevice, d.streams[b]));
            r = sponge_dev(ctx, (int)g, dev - e0 * 4, dev + elems_u64, cnt, dev + elems_u64 + off_u64, tag, d.streams[b]);
            if (r) return r;
            CUDA_TRY(ctx, cudaMemcpyAsync(h + elems_u64 + off_u64, dev + elems_u64 + off_u64, cnt * 32, cudaMemcpyDeviceToHost, d.streams[b]));
            CUDA_TRY(ctx, cudaEventRecord(d.done[b], d.streams[b]));
        }
        for (size_t c = n_chunks > (size_t)kNumBuf ? n_chunks - kNumBuf : 0; c < n_chunks; c++) {
            r = drain(c);
            if (r) return r;
        }
        return HADES_OK;
    };
    // chunk boundaries of every device range
    std::vector<std::vector<size_t>> cuts(G);
    for (size_t g = 0; g < G; g++) {
        cuts[g].push_back(bound[g]);
        size_t start = bound[g];
        for (size_t m = bound[g]; m < bound[g + 1]; m++) {
            const uint64_t len = offsets[m + 1] - offsets[m];
            if (len > kChunkElems) { cuts[g].assign({bound[g], bound[g + 1]}); break; }  // a giant message: single shot
            if (m > start && (offsets[m + 1] - offsets[start] > kChunkElems || m - start >= kChunkMsgs)) {
                cuts[g].push_back(m);
                start = m;
            }
        }
        if (cuts[g].back() != bound[g + 1]) cuts[g].push_back(bound[g + 1]);
    }
    std::vector<int> rcs(G, HADES_OK);
    std::vector<std::thread> th;
    auto run = [&](size_t g) {
        cudaSetDevice(ctx->devs[g].ordinal);
        rcs[g] = cuts[g].size() > 2 ? pipelined(g, cuts[g]) : single_shot(g, bound[g], bound[g + 1] - bound[g]);
    };
    for (size_t g = 1; g < G; g++) {
        if (cuts[g].size() > 2) th.emplace_back(run, g);  // a pipelined range blocks on its staging copies: own thread
        else run(g);
    }
    run(0);
    for (auto& t : th) t.join();
    for (int r : rcs) if (r != HADES_OK && rc == HADES_OK) rc = r;
    for (size_t g = 0; g < G; g++) {
        cudaSetDevice(ctx->devs[g].ordinal);
        for (int b = 0; b < kNumBuf; b++) {
            cudaError_t e = cudaStreamSynchronize(ctx->devs[g].streams[b]);
            if (e != cudaSuccess && rc == HADES_OK) rc = fail(ctx, HADES_ERR_CUDA, "sponge pass failed: %s", cudaGetErrorString(e));
        }
    }
    return rc;
}

int hades_sponge_batch(hades_ctx* ctx, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs, uint64_t* out) {
    return sponge_host(ctx, elems, offsets, n_msgs, out, nullptr);
}
int hades_sponge_batch_ds(hades_ctx* ctx, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs,
                          const uint64_t domain_tag[4], uint64_t* out) {
    if (!domain_tag) return fail(ctx, HADES_ERR_INVALID_ARG, "null domain tag");
    return sponge_host(ctx, elems, offsets, n_msgs, out, domain_tag);
}

int hades_host_register(hades_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    DeviceGuard guard;
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
    DeviceGuard guard;
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
    DeviceGuard guard;
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
    DeviceGuard guard;
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

int hades_fr_op_shape(int op, int* in_words, int* out_words) {
    if (!in_words || !out_words) return HADES_ERR_INVALID_ARG;
    return fr_test_shape(op, in_words, out_words) ? HADES_ERR_INVALID_ARG : HADES_OK;
}

int hades_fr_op_dev(hades_ctx* ctx, int dev_index, int op, const uint32_t* d_in, uint32_t* d_out, size_t n, void* stream) {
    if (!valid_dev(ctx, dev_index)) return fail(ctx, HADES_ERR_INVALID_ARG, "bad context or device index");
    int iw = 0, ow = 0;
    if (fr_test_shape(op, &iw, &ow)) return fail(ctx, HADES_ERR_INVALID_ARG, "unknown field operation %d", op);
    if (n && (!d_in || !d_out)) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    DeviceGuard guard;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[dev_index].ordinal));
    ctx->launches++;
    CUDA_TRY(ctx, fr_test_launch(op, d_in, d_out, n, (cudaStream_t)stream));
    return HADES_OK;
}

int hades_kernel_info(hades_ctx* ctx, const char* kernel, int* regs_per_thread, int* local_bytes,
                      int* max_threads_per_block) {
    if (!ctx || !kernel) return fail(ctx, HADES_ERR_INVALID_ARG, "null pointer");
    DeviceGuard guard;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    cudaFuncAttributes a;
    cudaError_t e = ctx->generic() ? (strcmp(kernel, "perm") ? cudaErrorInvalidValue : generic_func_attributes((int)ctx->width, &a))
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
    if (algo == 1 && !ctx->has_sparse) return fail(ctx, HADES_ERR_CONSTANTS, "the sparse schedule could not be derived for these constants");
    Variant probe = ctx->variant;
    probe.algo = algo;
    probe.regs = regs;
    if (!ctx->ops2[algo]->supports(probe))
        return fail(ctx, HADES_ERR_INVALID_ARG, "launch shape regs=%d is not built for width %u, algo %d (include/hades_cuda.h)", regs, ctx->width, algo);
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

int hades_set_coop_wide_threshold(hades_ctx* ctx, size_t max_states) {
    if (!ctx) return fail(ctx, HADES_ERR_INVALID_ARG, "null context");
    if (ctx->width != 5 || !ctx->has_ccf)
        return fail(ctx, HADES_ERR_INVALID_ARG, "the cooperative kernels exist for width 5 with the canonical-form tables only");
    ctx->variant.coop_wide_max = (int)std::min<size_t>(max_states, (size_t)1 << 24);
    return HADES_OK;
}

}  // extern "C"
