// Kernels for ONE permutation width (HADES_W), included by hades_w{3,5,9}.cu.  One thread owns one
// width-W state (or one Merkle node / one sponge message) and keeps it in registers across all 67
// rounds.  Memory traffic is 2*32*W bytes per permutation (320 B at W=5) against ~1e5 integer
// multiplies, so these kernels are bound by the integer-multiply pipe, not by HBM.
#pragma once
#if !defined(HADES_W) || !defined(HADES_ALGO)
#error "define HADES_W (3|5|9) and HADES_ALGO (0 dense | 1 sparse | 2 canonical form) before including width_impl.cuh"
#endif
#include <string.h>

#include "hades.cuh"
#if HADES_W == 5 && HADES_ALGO == 2
#include "coop.cuh"
#endif
#include "width_ops.hpp"

namespace hades {
namespace {

constexpr int W = HADES_W;
typedef OptLayout<W> Layout;
constexpr int kDenseEntries = kRounds * W + W * W;

// ---- constant tables of this width (uploaded once per device by hades_init) ----------------------
// dense: ROUND_CONSTANTS[0 .. 67W) (src/round_constants.rs:29-48) then MDS_MATRIX row-major
//        (src/mds_matrix.rs:18-40); opt: the derived layout of host_tables.hpp.  8 u32 limbs each.
// One table per translation unit: the dense and the optimised kernels of a width live in separate TUs
// (HADES_ALGO) because together their tables exceed the 64 KB constant bank.
constexpr int kAlgo = HADES_ALGO;
constexpr int kTableEntries = kAlgo == 0 ? kDenseEntries : kAlgo == 1 ? Layout::kEntries : CcfLayout<W>::kEntries;
__constant__ uint32_t c_table[kTableEntries * 8];
static_assert(sizeof(uint32_t) * kTableEntries * 8 + 32 <= 65536, "constant bank overflow");

struct DenseConsts {
    static __device__ __forceinline__ uint32_t ark(int idx, int k) { return c_table[idx * 8 + k]; }
    static __device__ __forceinline__ uint32_t mds(int r, int c, int k) { return c_table[(kRounds * W + r * W + c) * 8 + k]; }
};
struct OptTab {
    static __device__ __forceinline__ uint32_t tab(int entry, int k) { return c_table[entry * 8 + k]; }
    static __device__ __forceinline__ const uint32_t* ptr(int entry) { return c_table + entry * 8; }
};

// the fast schedules (1: sparse partial rounds, 2: canonical form), with or without the per-round barrier
template <class Sync>
__device__ __forceinline__ void permute_fast(Fr (&s)[W]) {
    if constexpr (kAlgo == 2) hades_perm_ccf<W, OptTab, Sync>(s);
    else hades_perm_opt<W, OptTab, Sync>(s);
}

template <int ALGO>
__device__ __forceinline__ void permute(Fr (&s)[W]) {
    static_assert(ALGO == kAlgo, "this translation unit holds one algorithm");
    if constexpr (ALGO == 0) hades_perm<W, DenseConsts>(s);
    else permute_fast<NoSync>(s);
}

// 32-byte element <-> registers through two 128-bit accesses (pointers are 16-byte aligned).
__device__ __forceinline__ void fr_load(Fr& x, const uint4* p) {
    uint4 a = p[0], b = p[1];
    x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
    x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
}
__device__ __forceinline__ void fr_store(uint4* p, const Fr& x) {
    p[0] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
    p[1] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
}

// ---- perm_batch: `Strategy::perm` (strategies.rs:140) over n independent states, in place ---------
template <int ALGO, int MINB>
__global__ void __launch_bounds__(kPermThreads, MINB) perm_batch_kernel(uint4* __restrict__ states, size_t n) {
    size_t i = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    if (i >= n) return;
    uint4* p = states + i * (2 * W);
    Fr s[W];
#pragma unroll
    for (int j = 0; j < W; j++) fr_load(s[j], p + 2 * j);
    permute<ALGO>(s);
#pragma unroll
    for (int j = 0; j < W; j++) fr_store(p + 2 * j, s[j]);
}

#if HADES_ALGO >= 1
// Lockstep variant: BLOCK threads per block, one barrier per round; out-of-range threads compute on the
// last state and skip the store so that every thread reaches every barrier.
struct BlockSync {
    static __device__ __forceinline__ void sync() { __syncthreads(); }
};
// States are array-of-structs (32*W bytes each), so a per-thread access pattern touches 32 different
// sectors per warp request.  The lockstep kernel therefore moves a warp's 32 states with fully coalesced
// 128-bit accesses (lane l moves 16-byte chunks l, l+32, ... of the warp's contiguous 32*32*W bytes) and
// transposes through shared memory.  Row pitch 2W+1 chunks (176 B at W=5): bank-conflict-free for the
// per-thread 128-bit reads.
constexpr int kChunksPerState = 2 * W;
constexpr int kStagePitch = 2 * W + 1;

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) perm_batch_lockstep_kernel(uint4* __restrict__ states, size_t n) {
    constexpr bool kStaged = (size_t)BLOCK * kStagePitch * 16 <= 48 * 1024;
    __shared__ uint4 stage[kStaged ? BLOCK * kStagePitch : 1];
    const int lane = threadIdx.x & 31;
    const size_t warp_first = (size_t)blockIdx.x * BLOCK + (threadIdx.x & ~31);
    const int n_in_warp = warp_first >= n ? 0 : (n - warp_first < 32 ? (int)(n - warp_first) : 32);
    const bool live = lane < n_in_warp;
    Fr s[W];
    if constexpr (kStaged) {
        uint4* tile = stage + (threadIdx.x & ~31) * kStagePitch;
        uint4* gbase = states + warp_first * kChunksPerState;
#pragma unroll
        for (int k = 0; k < kChunksPerState; k++) {
            const int c = lane + 32 * k, st = c / kChunksPerState, off = c % kChunksPerState;
            if (st < n_in_warp) tile[st * kStagePitch + off] = gbase[c];
        }
        __syncwarp();
        // dead lanes recompute the warp's first state (or zeros in a fully dead warp) and store nothing
        const uint4* mine = tile + (live ? lane : 0) * kStagePitch;
#pragma unroll
        for (int j = 0; j < W; j++) {
            if (n_in_warp > 0) fr_load(s[j], mine + 2 * j);
            else {
#pragma unroll
                for (int k = 0; k < 8; k++) s[j].l[k] = 0;
            }
        }
        __syncwarp();
        permute_fast<BlockSync>(s);
        // The indices are RECOMPUTED from the special registers (volatile reads, so they are not merged with
        // the ones above): nothing but the state is live across the permutation, which keeps the kernel
        // inside its register budget without a stack slot.
        unsigned tid2, bid2;
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid2));
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid2));
        const int lane2 = tid2 & 31;
        const size_t warp_first2 = (size_t)bid2 * BLOCK + (tid2 & ~31u);
        const int n_in_warp2 = warp_first2 >= n ? 0 : (n - warp_first2 < 32 ? (int)(n - warp_first2) : 32);
        uint4* tile2 = stage + (tid2 & ~31u) * kStagePitch;
        uint4* gbase2 = states + warp_first2 * kChunksPerState;
        if (lane2 < n_in_warp2) {
#pragma unroll
            for (int j = 0; j < W; j++) fr_store(tile2 + lane2 * kStagePitch + 2 * j, s[j]);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kChunksPerState; k++) {
            const int c = lane2 + 32 * k, st = c / kChunksPerState, off = c % kChunksPerState;
            if (st < n_in_warp2) gbase2[c] = tile2[st * kStagePitch + off];
        }
    } else {
        const size_t i = warp_first + lane;
        uint4* p = states + (i < n ? i : n - 1) * kChunksPerState;
#pragma unroll
        for (int j = 0; j < W; j++) fr_load(s[j], p + 2 * j);
        permute_fast<BlockSync>(s);
        if (i < n) {
#pragma unroll
            for (int j = 0; j < W; j++) fr_store(p + 2 * j, s[j]);
        }
    }
}

#endif  // HADES_ALGO >= 1

#if HADES_W == 5
// Montgomery forms of the two small constants the compositions need (checked in tests).
__device__ __forceinline__ void fr_set_one(Fr& x) {  // 1 * R mod p
    const uint32_t v[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                           0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
#pragma unroll
    for (int k = 0; k < 8; k++) x.l[k] = v[k];
}
__device__ __forceinline__ void fr_set_fifteen(Fr& x) {  // 15 * R mod p (Merkle bitmask 0b1111)
    const uint32_t v[8] = {0xffffffdfu, 0x00000020u, 0x00362421u, 0x348ddb9du,
                           0xc2232750u, 0x658b26f6u, 0xa2b2d9b1u, 0x0e5d6e47u};
#pragma unroll
    for (int k = 0; k < 8; k++) x.l[k] = v[k];
}
// Montgomery forms of the Merkle bitmasks 2^k - 1 for k = 1..4 present children (ragged trees; checked in tests)
__device__ const uint32_t kMaskMont[4][8] = {
    {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u},   // 1
    {0xfffffffau, 0x00000005u, 0x0009d806u, 0x098e27eeu, 0xc634efe0u, 0xcca4efcfu, 0x064f104eu, 0x486e140du},   // 3
    {0xfffffff1u, 0x0000000eu, 0x00189c0fu, 0x17e363d3u, 0x6f8457b0u, 0xff9c5787u, 0x8fc5a8c4u, 0x35133220u},   // 7
    {0xffffffdfu, 0x00000020u, 0x00362421u, 0x348ddb9du, 0xc2232750u, 0x658b26f6u, 0xa2b2d9b1u, 0x0e5d6e47u}};  // 15
// word 0 of a Merkle node with `present` (1..4) children; only the last node of a level can have fewer than 4
__device__ __forceinline__ void fr_set_mask(Fr& x, int present) {
    fr_set_fifteen(x);
    if (present < 4) {
#pragma unroll
        for (int k = 0; k < 8; k++) x.l[k] = kMaskMont[present > 0 ? present - 1 : 0][k];
    }
}
__device__ __forceinline__ void fr_set_zero(Fr& x) {
#pragma unroll
    for (int k = 0; k < 8; k++) x.l[k] = 0;
}

// ---- merkle level: out[i] = perm([2^k - 1, in[4i], ..., in[4i+k-1], 0 ...])[1], k = children present (4 except
// possibly for the last node of a ragged level: n_in = number of nodes of the input level, n_out = ceil(n_in / 4))
template <int ALGO, int MINB>
__global__ void __launch_bounds__(kPermThreads, MINB)
merkle_level_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_out, size_t n_in) {
    size_t i = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    if (i >= n_out) return;
    const int present = n_in - 4 * i < 4 ? (int)(n_in - 4 * i) : 4;
    Fr s[5];
    fr_set_mask(s[0], present);
    const uint4* p = in + i * 8;  // 4 children = 128 contiguous bytes
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (j < present) fr_load(s[1 + j], p + 2 * j);
        else fr_set_zero(s[1 + j]);
    }
    permute<ALGO>(s);
    fr_store(out + i * 2, s[1]);
}

#if HADES_ALGO >= 1
// Lockstep Merkle level: same launch shape as the perm kernel (one barrier per round, no early exit); a
// warp's 32 x 128 B of children are moved with coalesced 128-bit loads through padded shared memory.
template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
merkle_level_lockstep_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_out, size_t n_in) {
    constexpr int kPitch = 9;  // 8 chunks of children + 1 pad: conflict-free 128-bit reads
    __shared__ uint4 stage[BLOCK * kPitch];
    const int lane = threadIdx.x & 31;
    const size_t warp_first = (size_t)blockIdx.x * BLOCK + (threadIdx.x & ~31);
    const int n_in_warp = warp_first >= n_out ? 0 : (n_out - warp_first < 32 ? (int)(n_out - warp_first) : 32);
    const bool live = lane < n_in_warp;
    uint4* tile = stage + (threadIdx.x & ~31) * kPitch;
    const uint4* gbase = in + warp_first * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = lane + 32 * k, st = c >> 3, off = c & 7;
        // chunk c = half of child (off >> 1) of node warp_first + st; a ragged level ends inside a node
        if (st < n_in_warp && 4 * (warp_first + st) + (off >> 1) < n_in) tile[st * kPitch + off] = gbase[c];
    }
    __syncwarp();
    Fr s[5];
    const int my = live ? lane : 0;  // dead lanes recompute the warp's first node and store nothing
    const size_t first_child = 4 * (warp_first + my);
    const int present = n_in_warp == 0 ? 0 : (n_in - first_child < 4 ? (int)(n_in - first_child) : 4);
    fr_set_mask(s[0], present);
    const uint4* mine = tile + my * kPitch;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (j < present) fr_load(s[1 + j], mine + 2 * j);
        else fr_set_zero(s[1 + j]);
    }
    permute_fast<BlockSync>(s);
    // indices recomputed from the special registers (see perm_batch_lockstep_kernel): no stack slot
    unsigned tid2, bid2;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid2));
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid2));
    const size_t node = (size_t)bid2 * BLOCK + tid2;
    if (node < n_out) fr_store(out + node * 2, s[1]);
}
#endif  // HADES_ALGO >= 1

// ---- batched verification of Merkle openings: the consumer of hades_merkle_open_dev's branches --------------
// One thread per opening walks its path: at every level the stored group of four children must hold the running
// node at the path position and zeros where the (ragged) level has no child; the parent is
// perm([2^k - 1, c_0 .. c_{k-1}, 0 ...])[1]; after the last level the node must equal the root.
// (Build-defined like the tree itself; the checker restates it as `merkle_verify`.)
__device__ __forceinline__ bool fr_equal(const Fr& a, const Fr& b) {
    uint32_t d = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) d |= a.l[k] ^ b.l[k];
    return d == 0;
}
__device__ __forceinline__ bool fr_is_zero(const Fr& a) {
    uint32_t d = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) d |= a.l[k];
    return d == 0;
}
template <int ALGO, int MINB>
__global__ void __launch_bounds__(kPermThreads, MINB)
merkle_verify_kernel(const uint4* __restrict__ leaves, const uint64_t* __restrict__ index, size_t n_open, size_t n_leaves,
                     int levels, const uint4* __restrict__ branch, const uint4* __restrict__ root, uint32_t* __restrict__ ok) {
    const size_t o = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    if (o >= n_open) return;
    size_t i = index[o], m = n_leaves;
    bool good = i < n_leaves;
    if (!good) i = 0;  // keep every address in range; the verdict is already false
    Fr node;
    fr_load(node, leaves + i * 2);
    const uint4* b = branch + o * (size_t)levels * 8;
#pragma unroll 1
    for (int l = 0; l < levels; l++) {
        const size_t first = 4 * (i / 4);
        const int k = m - first < 4 ? (int)(m - first) : 4;
        const int pos = (int)(i & 3);
        Fr s[5];
        fr_set_mask(s[0], k);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            fr_load(s[1 + j], b + l * 8 + 2 * j);
            if (j == pos) good &= fr_equal(s[1 + j], node);
            if (j >= k) good &= fr_is_zero(s[1 + j]);
        }
        permute<ALGO>(s);
        node = s[1];
        i >>= 2;
        m = (m + 3) / 4;
    }
    Fr r;
    fr_load(r, root);
    good &= (m == 1) && fr_equal(node, r);
    ok[o] = good ? 1u : 0u;
}

// ---- sponge: rate 4 / capacity 1, one message per thread (CSR offsets) ------------------------------
// `order` (optional) maps thread -> message so that a warp works on messages of equal block count.
template <int ALGO, int MINB>
__global__ void __launch_bounds__(kPermThreads, MINB)
sponge_kernel(const uint4* __restrict__ elems, const uint64_t* __restrict__ offsets,
              const uint32_t* __restrict__ order, uint4* __restrict__ out, size_t n_threads, const SpongeTag tag) {
    size_t t = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    if (t >= n_threads) return;
    size_t m = order ? order[t] : t;
    uint64_t b = offsets[m], e = offsets[m + 1];
    Fr s[5];
#pragma unroll
    for (int j = 0; j < 5; j++) fr_set_zero(s[j]);
#pragma unroll
    for (int k = 0; k < 8; k++) s[0].l[k] = tag.l[k];  // capacity word: zero or the domain tag
    bool padded = false;
#pragma unroll 1
    while (!padded) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            Fr x;
            bool add = true;
            if (b < e) {
                fr_load(x, elems + b * 2);
                b++;
            } else if (!padded) {
                fr_set_one(x);
                padded = true;
            } else {
                add = false;
            }
            if (add) fr_add(s[1 + k], s[1 + k], x);
        }
        permute<ALGO>(s);
    }
    fr_store(out + m * 2, s[1]);
}
#if HADES_ALGO >= 1
// Lockstep sponge: messages arrive sorted by permutation count (`order`), so the 128 messages of a block
// almost always need the same number of perms.  Every thread runs the block's MAXIMUM count (one barrier per
// round keeps the block on the same instruction-cache lines); a thread whose message ended earlier keeps
// permuting a dead state and has already captured its digest.
__global__ void __launch_bounds__(kPermThreads, 4)
sponge_lockstep_kernel(const uint4* __restrict__ elems, const uint64_t* __restrict__ offsets,
                       const uint32_t* __restrict__ order, uint4* __restrict__ out, size_t n_threads, const SpongeTag tag) {
    __shared__ unsigned int s_max_blocks;
    if (threadIdx.x == 0) s_max_blocks = 0;
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * kPermThreads + threadIdx.x;
    const bool live = t < n_threads;
    const size_t m = live ? (order ? order[t] : t) : 0;
    uint64_t b = live ? offsets[m] : 0, e = live ? offsets[m + 1] : 0;
    const unsigned int my_blocks = live ? (unsigned int)((e - b) / 4 + 1) : 0u;
    atomicMax(&s_max_blocks, my_blocks);
    __syncthreads();
    const unsigned int trips = s_max_blocks;
    Fr s[5], digest;
#pragma unroll
    for (int j = 0; j < 5; j++) fr_set_zero(s[j]);
#pragma unroll
    for (int k = 0; k < 8; k++) s[0].l[k] = tag.l[k];  // capacity word: zero or the domain tag
    fr_set_zero(digest);
    bool padded = false;
#pragma unroll 1
    for (unsigned int trip = 0; trip < trips; trip++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            Fr x;
            bool add = true;
            if (b < e) {
                fr_load(x, elems + b * 2);
                b++;
            } else if (!padded) {
                fr_set_one(x);
                padded = true;
            } else {
                add = false;
            }
            if (add) fr_add(s[1 + k], s[1 + k], x);
        }
        permute_fast<BlockSync>(s);
        if (trip + 1 == my_blocks) digest = s[1];
    }
    if (live) fr_store(out + m * 2, digest);
}
#endif  // HADES_ALGO >= 1
#endif  // HADES_W == 5


#if HADES_W == 5 && HADES_ALGO == 2
// ---- cooperative small-batch kernels (coop.cuh): one state per 8 lanes, table staged in shared memory ---------
// The lanes of a group read DIFFERENT table entries in the same instruction, which the constant cache would
// serialise; the 24 KB canonical-form table is therefore mirrored in global memory (uploaded with the constant
// bank) and copied to shared memory by every block.
__device__ uint4 g_coop_table[kTableEntries * 2];
extern __shared__ __align__(16) uint32_t coop_smem[];
struct CoopTab {
    static __device__ __forceinline__ uint32_t tab(int entry, int k) { return coop_smem[entry * 8 + k]; }
    static __device__ __forceinline__ const uint4* ptr4(int entry) { return reinterpret_cast<const uint4*>(coop_smem) + entry * 2; }
};
constexpr int kCoopBlock = 128;                        // 16 states per block, one warp per scheduler
constexpr int kCoopStatesPerBlock = kCoopBlock / kCoopLanes;
constexpr size_t kCoopSmemBytes = (size_t)kTableEntries * 32;
static_assert(kCoopSmemBytes <= 48 * 1024, "table must fit the default dynamic shared memory limit");

__device__ __forceinline__ void coop_stage_table() {
    uint4* dst = reinterpret_cast<uint4*>(coop_smem);
    for (int i = threadIdx.x; i < kTableEntries * 2; i += kCoopBlock) dst[i] = g_coop_table[i];
    __syncthreads();
}

// G = lanes per state: kCoopLanes (8) or kCoopWide (a warp; lowest latency, a quarter of the states per block)
template <int G>
__global__ void __launch_bounds__(kCoopBlock) perm_batch_coop_kernel(uint4* __restrict__ states, size_t n) {
    coop_stage_table();
    const int lane = threadIdx.x & (G - 1);
    const size_t g = (size_t)blockIdx.x * (kCoopBlock / G) + (threadIdx.x / G);
    const bool live = g < n;
    Fr s[5];
    const uint4* p = states + (live ? g : 0) * 10;  // all lanes of the group read the same 160 bytes (broadcast)
#pragma unroll
    for (int j = 0; j < 5; j++) {
        if (live) fr_load(s[j], p + 2 * j);
        else fr_set_zero(s[j]);
    }
    hades_perm_coop<CoopTab, G>(s, lane);
    if (live && lane < 5) {  // lane j writes word j
        Fr w;
        coop_select_word<5>(w.l, s, lane);
        fr_store(states + g * 10 + 2 * lane, w);
    }
}

template <int G>
__global__ void __launch_bounds__(kCoopBlock)
merkle_level_coop_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_out, size_t n_in) {
    coop_stage_table();
    const int lane = threadIdx.x & (G - 1);
    const size_t g = (size_t)blockIdx.x * (kCoopBlock / G) + (threadIdx.x / G);
    const bool live = g < n_out;
    const size_t first_child = 4 * (live ? g : 0);
    const int present = !live ? 0 : (n_in - first_child < 4 ? (int)(n_in - first_child) : 4);
    Fr s[5];
    fr_set_mask(s[0], present);
    const uint4* p = in + first_child * 2;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (j < present) fr_load(s[1 + j], p + 2 * j);
        else fr_set_zero(s[1 + j]);
    }
    hades_perm_coop<CoopTab, G>(s, lane);
    if (live && lane == 0) fr_store(out + g * 2, s[1]);
}

// Sponge for few messages (a lone hash is the commonest call a Poseidon user makes): one message per 8-lane group,
// no length bucketing; the four groups of a warp run the warp's maximum block count (all 32 lanes must reach every
// shuffle), a group whose message ended earlier keeps permuting a dead state and has already captured its digest.
template <int G>
__global__ void __launch_bounds__(kCoopBlock)
sponge_coop_kernel(const uint4* __restrict__ elems, const uint64_t* __restrict__ offsets, uint4* __restrict__ out, size_t n_msgs,
                   const SpongeTag tag) {
    coop_stage_table();
    const int lane = threadIdx.x & (G - 1);
    const size_t g = (size_t)blockIdx.x * (kCoopBlock / G) + (threadIdx.x / G);
    const bool live = g < n_msgs;
    uint64_t b = live ? offsets[g] : 0, e = live ? offsets[g + 1] : 0;
    const unsigned my_blocks = live ? (unsigned)((e - b) / 4 + 1) : 0u;
    unsigned trips = my_blocks;
    if constexpr (G < 32) {  // several groups per warp: all of them run the warp's maximum
        trips = max(trips, __shfl_xor_sync(0xffffffffu, trips, 8));
        trips = max(trips, __shfl_xor_sync(0xffffffffu, trips, 16));
    }
    Fr s[5], digest;
#pragma unroll
    for (int j = 0; j < 5; j++) fr_set_zero(s[j]);
#pragma unroll
    for (int k = 0; k < 8; k++) s[0].l[k] = tag.l[k];
    fr_set_zero(digest);
    bool padded = false;
#pragma unroll 1
    for (unsigned trip = 0; trip < trips; trip++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            Fr x;
            bool add = true;
            if (b < e) {
                fr_load(x, elems + b * 2);  // the 8 lanes of the group read the same element (broadcast)
                b++;
            } else if (!padded) {
                fr_set_one(x);
                padded = true;
            } else {
                add = false;
            }
            if (add) fr_add(s[1 + k], s[1 + k], x);
        }
        hades_perm_coop<CoopTab, G>(s, lane);
        if (trip + 1 == my_blocks) digest = s[1];
    }
    if (live && lane == 0) fr_store(out + g * 2, digest);
}

// Verification of few openings (a lone verify is a chain of `levels` permutations): one opening per 8-lane group.
template <int G>
__global__ void __launch_bounds__(kCoopBlock)
merkle_verify_coop_kernel(const uint4* __restrict__ leaves, const uint64_t* __restrict__ index, size_t n_open, size_t n_leaves,
                          int levels, const uint4* __restrict__ branch, const uint4* __restrict__ root, uint32_t* __restrict__ ok) {
    coop_stage_table();
    const int lane = threadIdx.x & (G - 1);
    const size_t o = (size_t)blockIdx.x * (kCoopBlock / G) + (threadIdx.x / G);
    const bool live = o < n_open;
    size_t i = live ? index[o] : 0, m = n_leaves;
    bool good = live && i < n_leaves;
    if (!good) i = 0;
    Fr node;
    fr_load(node, leaves + i * 2);
    const uint4* b = branch + (live ? o : 0) * (size_t)levels * 8;
#pragma unroll 1
    for (int l = 0; l < levels; l++) {  // `levels` is uniform: every lane of the warp reaches every shuffle
        const size_t first = 4 * (i / 4);
        const int k = m - first < 4 ? (int)(m - first) : 4;
        const int pos = (int)(i & 3);
        Fr s[5];
        fr_set_mask(s[0], k);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            fr_load(s[1 + j], b + l * 8 + 2 * j);
            if (j == pos) good &= fr_equal(s[1 + j], node);
            if (j >= k) good &= fr_is_zero(s[1 + j]);
        }
        hades_perm_coop<CoopTab, G>(s, lane);
        node = s[1];
        i >>= 2;
        m = (m + 3) / 4;
    }
    Fr r;
    fr_load(r, root);
    good &= (m == 1) && fr_equal(node, r);
    if (live && lane == 0) ok[o] = good ? 1u : 0u;
}
#endif  // cooperative kernels

// ---- host-side launchers ---------------------------------------------------------------------------
cudaError_t upload(const uint64_t* table) {
    cudaError_t e = upload_modulus();
    if (e != cudaSuccess) return e;
#if HADES_W == 5 && HADES_ALGO == 2
    e = cudaMemcpyToSymbol(g_coop_table, table, sizeof(c_table), 0, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
#endif
    return cudaMemcpyToSymbol(c_table, table, sizeof(c_table), 0, cudaMemcpyHostToDevice);
}

// regs: 0 -> minBlocks 4 (<=128 registers), 1 -> 3 (<=168), 2 -> 2 (<=255), 3 -> 5 (<=96)
#define HADES_DISPATCH(KERNEL, v, ...)                                  \
    do {                                                                \
        switch ((v).regs) {                                             \
            case 0: KERNEL<kAlgo, 4> __VA_ARGS__; break;                \
            case 1: KERNEL<kAlgo, 3> __VA_ARGS__; break;                \
            case 2: KERNEL<kAlgo, 2> __VA_ARGS__; break;                \
            default: KERNEL<kAlgo, 5> __VA_ARGS__; break;               \
        }                                                               \
    } while (0)

cudaError_t launch_perm(Variant v, uint64_t* d_states, size_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
#if HADES_W == 5 && HADES_ALGO == 2
    if (n <= (size_t)v.coop_max) {  // small batch: 8 lanes, or a whole warp, per state (latency kernels)
        if (n <= (size_t)v.coop_wide_max) {
            constexpr int kPer = kCoopBlock / kCoopWide;
            perm_batch_coop_kernel<kCoopWide><<<(unsigned)((n + kPer - 1) / kPer), kCoopBlock, kCoopSmemBytes, s>>>(
                reinterpret_cast<uint4*>(d_states), n);
            return cudaGetLastError();
        }
        const unsigned blocks = (unsigned)((n + kCoopStatesPerBlock - 1) / kCoopStatesPerBlock);
        perm_batch_coop_kernel<kCoopLanes><<<blocks, kCoopBlock, kCoopSmemBytes, s>>>(reinterpret_cast<uint4*>(d_states), n);
        return cudaGetLastError();
    }
#endif
#if HADES_ALGO >= 1
    if (v.regs >= 4) {  // lockstep launches: regs 4 -> 256 threads per block, 5 -> 512, 6.. experimental
        uint4* p = reinterpret_cast<uint4*>(d_states);
        switch (v.regs) {
            case 4: perm_batch_lockstep_kernel<256, 2><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n); break;
            case 5: perm_batch_lockstep_kernel<512, 1><<<(unsigned)((n + 511) / 512), 512, 0, s>>>(p, n); break;
#if HADES_W == 9
            case 6: perm_batch_lockstep_kernel<128, 2><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, n); break;
            case 7: perm_batch_lockstep_kernel<128, 3><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, n); break;
#endif
#if HADES_W == 3
            case 6: perm_batch_lockstep_kernel<128, 7><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, n); break;
            case 7: perm_batch_lockstep_kernel<128, 5><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, n); break;
#endif
#if HADES_W == 5
            case 6: perm_batch_lockstep_kernel<128, 5><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, n); break;
            case 7: perm_batch_lockstep_kernel<384, 1><<<(unsigned)((n + 383) / 384), 384, 0, s>>>(p, n); break;
            case 8: perm_batch_lockstep_kernel<640, 1><<<(unsigned)((n + 639) / 640), 640, 0, s>>>(p, n); break;
            case 9: perm_batch_lockstep_kernel<128, 4><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, n); break;
            case 10: perm_batch_lockstep_kernel<128, 6><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, n); break;
#endif
            default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
#endif
    size_t blocks = (n + kPermThreads - 1) / kPermThreads;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    if (v.regs >= 4) v.regs = 0;  // dense schedule has no lockstep build
    HADES_DISPATCH(perm_batch_kernel, v, <<<(unsigned)blocks, kPermThreads, 0, s>>>(reinterpret_cast<uint4*>(d_states), n));
    return cudaGetLastError();
}

#if HADES_W == 5
cudaError_t launch_merkle_level(Variant v, const uint64_t* d_in, uint64_t* d_out, size_t n_out, size_t n_in, cudaStream_t s) {
    if (n_out == 0) return cudaSuccess;
    if (n_in > 4 * n_out || n_in + 3 < 4 * n_out) return cudaErrorInvalidValue;  // n_out == ceil(n_in / 4)
#if HADES_ALGO == 2
    if (n_out <= (size_t)v.coop_max) {  // small level: 8 lanes, or a whole warp, per node (latency kernels)
        if (n_out <= (size_t)v.coop_wide_max) {
            constexpr int kPer = kCoopBlock / kCoopWide;
            merkle_level_coop_kernel<kCoopWide><<<(unsigned)((n_out + kPer - 1) / kPer), kCoopBlock, kCoopSmemBytes, s>>>(
                reinterpret_cast<const uint4*>(d_in), reinterpret_cast<uint4*>(d_out), n_out, n_in);
            return cudaGetLastError();
        }
        const unsigned blocks = (unsigned)((n_out + kCoopStatesPerBlock - 1) / kCoopStatesPerBlock);
        merkle_level_coop_kernel<kCoopLanes><<<blocks, kCoopBlock, kCoopSmemBytes, s>>>(
            reinterpret_cast<const uint4*>(d_in), reinterpret_cast<uint4*>(d_out), n_out, n_in);
        return cudaGetLastError();
    }
#endif
#if HADES_ALGO >= 1
    if (v.regs >= 4) {  // lockstep launch shapes share one Merkle build
        merkle_level_lockstep_kernel<128, 5><<<(unsigned)((n_out + 127) / 128), 128, 0, s>>>(
            reinterpret_cast<const uint4*>(d_in), reinterpret_cast<uint4*>(d_out), n_out, n_in);
        return cudaGetLastError();
    }
#endif
    if (v.regs >= 4) v.regs = 0;
    size_t blocks = (n_out + kPermThreads - 1) / kPermThreads;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    HADES_DISPATCH(merkle_level_kernel, v,
                   <<<(unsigned)blocks, kPermThreads, 0, s>>>(reinterpret_cast<const uint4*>(d_in), reinterpret_cast<uint4*>(d_out), n_out, n_in));
    return cudaGetLastError();
}
cudaError_t launch_merkle_verify(Variant v, const uint64_t* d_leaves, const uint64_t* d_index, size_t n_open, size_t n_leaves, int levels,
                                 const uint64_t* d_branch, const uint64_t* d_root, uint32_t* d_ok, cudaStream_t s) {
    if (n_open == 0) return cudaSuccess;
#if HADES_ALGO == 2
    if (n_open <= (size_t)v.coop_max) {  // few openings: 8 lanes, or a warp, per opening (latency kernels)
        if (n_open <= (size_t)v.coop_wide_max) {
            constexpr int kPer = kCoopBlock / kCoopWide;
            merkle_verify_coop_kernel<kCoopWide><<<(unsigned)((n_open + kPer - 1) / kPer), kCoopBlock, kCoopSmemBytes, s>>>(
                reinterpret_cast<const uint4*>(d_leaves), d_index, n_open, n_leaves, levels, reinterpret_cast<const uint4*>(d_branch),
                reinterpret_cast<const uint4*>(d_root), d_ok);
            return cudaGetLastError();
        }
        const unsigned cblocks = (unsigned)((n_open + kCoopStatesPerBlock - 1) / kCoopStatesPerBlock);
        merkle_verify_coop_kernel<kCoopLanes><<<cblocks, kCoopBlock, kCoopSmemBytes, s>>>(
            reinterpret_cast<const uint4*>(d_leaves), d_index, n_open, n_leaves, levels, reinterpret_cast<const uint4*>(d_branch),
            reinterpret_cast<const uint4*>(d_root), d_ok);
        return cudaGetLastError();
    }
#else
    (void)v;
#endif
    const size_t blocks = (n_open + kPermThreads - 1) / kPermThreads;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    merkle_verify_kernel<kAlgo, 4><<<(unsigned)blocks, kPermThreads, 0, s>>>(
        reinterpret_cast<const uint4*>(d_leaves), d_index, n_open, n_leaves, levels, reinterpret_cast<const uint4*>(d_branch),
        reinterpret_cast<const uint4*>(d_root), d_ok);
    return cudaGetLastError();
}
cudaError_t launch_sponge(Variant v, const uint64_t* d_elems, const uint64_t* d_offsets, const uint32_t* d_order,
                          uint64_t* d_out, size_t n_threads, SpongeTag tag, cudaStream_t s) {
    if (n_threads == 0) return cudaSuccess;
#if HADES_ALGO == 2
    if (n_threads <= (size_t)v.coop_max) {  // few messages: 8 lanes, or a warp, per message (latency kernels), no bucketing needed
        if (n_threads <= (size_t)v.coop_wide_max) {
            constexpr int kPer = kCoopBlock / kCoopWide;
            sponge_coop_kernel<kCoopWide><<<(unsigned)((n_threads + kPer - 1) / kPer), kCoopBlock, kCoopSmemBytes, s>>>(
                reinterpret_cast<const uint4*>(d_elems), d_offsets, reinterpret_cast<uint4*>(d_out), n_threads, tag);
            return cudaGetLastError();
        }
        const unsigned blocks = (unsigned)((n_threads + kCoopStatesPerBlock - 1) / kCoopStatesPerBlock);
        sponge_coop_kernel<kCoopLanes><<<blocks, kCoopBlock, kCoopSmemBytes, s>>>(
            reinterpret_cast<const uint4*>(d_elems), d_offsets, reinterpret_cast<uint4*>(d_out), n_threads, tag);
        return cudaGetLastError();
    }
#endif
    // lockstep shapes do not apply (messages differ in length); the optimised kernel fits 96 registers, so
    // use 5 blocks/SM (measured: 225 ms vs 249 ms at 4 blocks/SM for the 2^22-message config)
#if HADES_ALGO >= 1
    if (v.regs >= 4 && d_order != nullptr) {  // lockstep launch shapes: sorted messages, block-uniform trip counts
        sponge_lockstep_kernel<<<(unsigned)((n_threads + kPermThreads - 1) / kPermThreads), kPermThreads, 0, s>>>(
            reinterpret_cast<const uint4*>(d_elems), d_offsets, d_order, reinterpret_cast<uint4*>(d_out), n_threads, tag);
        return cudaGetLastError();
    }
#endif
    if (v.regs >= 4) v.regs = (kAlgo >= 1) ? 3 : 0;
    size_t blocks = (n_threads + kPermThreads - 1) / kPermThreads;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    HADES_DISPATCH(sponge_kernel, v,
                   <<<(unsigned)blocks, kPermThreads, 0, s>>>(reinterpret_cast<const uint4*>(d_elems), d_offsets, d_order,
                                                             reinterpret_cast<uint4*>(d_out), n_threads, tag));
    return cudaGetLastError();
}
#endif

#define HADES_ATTR(KERNEL, v, out)                                              \
    ((v).regs == 0   ? cudaFuncGetAttributes(out, KERNEL<kAlgo, 4>)             \
     : (v).regs == 1 ? cudaFuncGetAttributes(out, KERNEL<kAlgo, 3>)             \
     : (v).regs == 2 ? cudaFuncGetAttributes(out, KERNEL<kAlgo, 2>)             \
                     : cudaFuncGetAttributes(out, KERNEL<kAlgo, 5>))

cudaError_t func_attributes(const char* kernel, Variant v, cudaFuncAttributes* out) {
#if HADES_W == 5 && HADES_ALGO == 2
    if (!strcmp(kernel, "perm_coop")) return cudaFuncGetAttributes(out, perm_batch_coop_kernel<kCoopLanes>);
    if (!strcmp(kernel, "merkle_coop")) return cudaFuncGetAttributes(out, merkle_level_coop_kernel<kCoopLanes>);
    if (!strcmp(kernel, "perm_coop_wide")) return cudaFuncGetAttributes(out, perm_batch_coop_kernel<kCoopWide>);
    if (!strcmp(kernel, "merkle_coop_wide")) return cudaFuncGetAttributes(out, merkle_level_coop_kernel<kCoopWide>);
    if (!strcmp(kernel, "sponge_coop")) return cudaFuncGetAttributes(out, sponge_coop_kernel<kCoopLanes>);
    if (!strcmp(kernel, "sponge_coop_wide")) return cudaFuncGetAttributes(out, sponge_coop_kernel<kCoopWide>);
#endif
#if HADES_ALGO >= 1
    if (!strcmp(kernel, "perm") && v.regs == 4) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<256, 2>);
    if (!strcmp(kernel, "perm") && v.regs == 5) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<512, 1>);
#if HADES_W == 9
    if (!strcmp(kernel, "perm") && v.regs == 6) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<128, 2>);
    if (!strcmp(kernel, "perm") && v.regs == 7) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<128, 3>);
#endif
#if HADES_W == 3
    if (!strcmp(kernel, "perm") && v.regs == 6) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<128, 7>);
    if (!strcmp(kernel, "perm") && v.regs == 7) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<128, 5>);
#endif
#if HADES_W == 5
    if (!strcmp(kernel, "perm") && v.regs == 6) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<128, 5>);
    if (!strcmp(kernel, "perm") && v.regs == 7) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<384, 1>);
    if (!strcmp(kernel, "perm") && v.regs == 8) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<640, 1>);
    if (!strcmp(kernel, "perm") && v.regs == 9) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<128, 4>);
    if (!strcmp(kernel, "perm") && v.regs == 10) return cudaFuncGetAttributes(out, perm_batch_lockstep_kernel<128, 6>);
#endif
#if HADES_W == 5
    if (!strcmp(kernel, "merkle") && v.regs >= 4) return cudaFuncGetAttributes(out, merkle_level_lockstep_kernel<128, 5>);
#endif
#endif  // HADES_ALGO >= 1
    if (v.regs >= 4) v.regs = 0;  // sponge / dense: plain 128-thread launches
    if (!strcmp(kernel, "perm")) return HADES_ATTR(perm_batch_kernel, v, out);
#if HADES_W == 5
    if (!strcmp(kernel, "merkle")) return HADES_ATTR(merkle_level_kernel, v, out);
    if (!strcmp(kernel, "merkle_verify")) return cudaFuncGetAttributes(out, merkle_verify_kernel<kAlgo, 4>);
    if (!strcmp(kernel, "sponge")) return HADES_ATTR(sponge_kernel, v, out);
#endif
    return cudaErrorInvalidValue;
}

// launch shapes built for this width / schedule (must mirror the switch in launch_perm)
bool supports(Variant v) {
    if (v.regs < 0) return false;
    if (v.regs <= 3) return true;
    if (kAlgo == 0) return false;  // the dense schedule has no lockstep build
    if (v.regs <= 5) return true;
    return v.regs <= (W == 5 ? 10 : 7);
}

const WidthOps kOps = {W, kAlgo, (size_t)kTableEntries * 4, upload, launch_perm,
#if HADES_W == 5
                       launch_merkle_level, launch_sponge, launch_merkle_verify,
#else
                       nullptr, nullptr, nullptr,
#endif
                       func_attributes, supports};

}  // namespace
}  // namespace hades
