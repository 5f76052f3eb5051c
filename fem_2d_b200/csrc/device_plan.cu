// Device part of the symbolic phase: the fixed upper-triangular pattern (BTreeMap<[u32;2]> order, sparse_matrix.rs:16,48-58) and the
// per-slot source map that replaces the reference's per-Elem BTreeMap inserts and serial merge (sparse_matrix.rs:68-120,
// linalg.rs:59-81).  Symbolic assembly by rows: a DoF that lives in exactly one pair block (an Elem-type DoF of an Elem without
// ancestor specs: 98 % of the slots at 1 M DoFs) gets its row written directly -- its columns are the tail of its Elem's sorted DoF
// list -- and only the pairs whose row DoF is shared between blocks (edge-type DoFs, RBS ancestor/descendant overlaps) go through
// keygen -> radix sort -> unique, which is also where keys with more than one contribution come from.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "device_plan.hpp"

namespace fem2d {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return FEM2D_ERR_CUDA; } } while (0)

namespace {

struct DevBlock {
    unsigned long long pair_off;
    unsigned long long v_off;
    uint32_t dofP_off, dofQ_off;   // into canon_dof
    uint32_t nP, nQ;
    uint32_t local, pad;
};

// Number of pair blocks every DoF takes part in (as a P function or as a Q function).  1 = the local block of its own Elem only.
__global__ void incidence_kernel(const DevBlock* __restrict__ blocks, const uint32_t* __restrict__ canon_dof, uint32_t* __restrict__ cnt) {
    const DevBlock b = blocks[blockIdx.x];
    for (uint32_t t = threadIdx.x; t < b.nP; t += blockDim.x) atomicAdd(&cnt[canon_dof[b.dofP_off + t]], 1u);
    if (!b.local) for (uint32_t t = threadIdx.x; t < b.nQ; t += blockDim.x) atomicAdd(&cnt[canon_dof[b.dofQ_off + t]], 1u);
}

constexpr uint32_t ORDER_MAX = 1024;   // >= the largest BasisSpec list (2 * 20 * 21 = 840 at MAX_POLYNOMIAL_ORDER)

// One CTA per LOCAL block: sorts the Elem's DoF ids (bitonic, shared memory), stores the canonical index of the k-th smallest DoF
// (perm) and, for every direct DoF (cnt == 1), the length of its row: the DoFs of the Elem that are >= it.
__global__ void __launch_bounds__(256) local_order_kernel(const DevBlock* __restrict__ blocks, const uint32_t* __restrict__ canon_dof,
                                                          const uint32_t* __restrict__ cnt, uint16_t* __restrict__ perm,
                                                          uint32_t* __restrict__ len_all, uint32_t* __restrict__ len_dir) {
    const DevBlock b = blocks[blockIdx.x];
    if (!b.local) return;
    __shared__ uint32_t s_key[ORDER_MAX];
    __shared__ uint16_t s_idx[ORDER_MAX];
    uint32_t n2 = 1; while (n2 < b.nP) n2 <<= 1;
    for (uint32_t t = threadIdx.x; t < n2; t += blockDim.x) { s_key[t] = t < b.nP ? canon_dof[b.dofP_off + t] : 0xffffffffu; s_idx[t] = (uint16_t)t; }
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < n2; t += blockDim.x) {
                const uint32_t x = t ^ j;
                if (x > t) {
                    const bool up = (t & k) == 0;
                    const uint32_t a = s_key[t], c = s_key[x];
                    if ((a > c) == up) { s_key[t] = c; s_key[x] = a; const uint16_t ia = s_idx[t]; s_idx[t] = s_idx[x]; s_idx[x] = ia; }
                }
            }
            __syncthreads();
        }
    for (uint32_t k = threadIdx.x; k < b.nP; k += blockDim.x) {
        perm[b.dofP_off + k] = s_idx[k];
        const uint32_t dof = s_key[k];
        if (cnt[dof] == 1u) { len_all[dof] = b.nP - k; len_dir[dof] = b.nP - k; }
    }
}

// Pairs of a block whose row DoF (the smaller one, sparse_matrix.rs:78-91) is shared between blocks, in the block's generation order
// (a-major, q ascending; local blocks: q >= a).  WRITE = false: count them; WRITE = true: emit key = [row << 32 | col] and the
// pair's index in V at sub_off[block] + rank (ordered compaction, so equal keys keep the reference's generation order).
template <bool WRITE>
__global__ void __launch_bounds__(256) shared_pairs_kernel(const DevBlock* __restrict__ blocks, const uint32_t* __restrict__ canon_dof,
                                                           const uint32_t* __restrict__ cnt, uint32_t* __restrict__ sub_cnt,
                                                           const uint32_t* __restrict__ sub_off, unsigned long long* __restrict__ keys,
                                                           uint32_t* __restrict__ srcs) {
    typedef cub::BlockScan<uint32_t, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ uint32_t s_run;
    const DevBlock b = blocks[blockIdx.x];
    const uint32_t* dp = canon_dof + b.dofP_off;
    const uint32_t* dq = canon_dof + b.dofQ_off;
    const uint32_t total = b.nP * b.nQ;
    if (threadIdx.x == 0) s_run = WRITE ? sub_off[blockIdx.x] : 0u;
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t t0 = 0; t0 < total; t0 += blockDim.x) {
        const uint32_t t = t0 + threadIdx.x;
        bool take = false; uint32_t r = 0, c = 0;
        if (t < total) {
            const uint32_t a = t / b.nQ, q = t - a * b.nQ;
            if (!b.local || q >= a) {
                const uint32_t x = dp[a], y = dq[q];
                r = x < y ? x : y; c = x < y ? y : x;
                take = cnt[r] != 1u;
            }
        }
        if (WRITE) {
            uint32_t rank, sum;
            Scan(tmp).ExclusiveSum(take ? 1u : 0u, rank, sum);
            const uint32_t base = s_run;
            if (take) { keys[base + rank] = (unsigned long long)r << 32 | c; srcs[base + rank] = (uint32_t)(b.v_off + t); }
            __syncthreads();
            if (threadIdx.x == 0) s_run = base + sum;
            __syncthreads();
        } else mine += take ? 1u : 0u;
    }
    if (!WRITE) {
        uint32_t sum;
        Scan(tmp).ExclusiveSum(mine, mine, sum);
        if (threadIdx.x == 0) sub_cnt[blockIdx.x] = sum;
    }
}

__global__ void head_flags_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ head, uint32_t* __restrict__ len_all) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool h = i == 0 || keys[i] != keys[i - 1];
    head[i] = h ? 1u : 0u;
    if (h) atomicAdd(&len_all[(uint32_t)(keys[i] >> 32)], 1u);   // row length of a shared row = its unique keys
}

// Sorted shared pairs -> pattern.  rank = (inclusive scan of head)[i] - 1 is the key's rank among the shared rows' slots; its slot is
// rank + dir_before[row] (the direct rows' slots that precede it).  Heads write the pattern, the others go to the extras list (kept in
// sorted order: position i - rank - 1 is the rank among non-heads).
__global__ void emit_shared_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ srcs,
                                   const uint32_t* __restrict__ head, const uint32_t* __restrict__ scan, uint32_t n,
                                   const uint32_t* __restrict__ dir_before,
                                   uint32_t* __restrict__ rows, uint32_t* __restrict__ cols, uint32_t* __restrict__ src1,
                                   uint32_t* __restrict__ extra_slot, uint32_t* __restrict__ extra_src, uint32_t* __restrict__ extra_first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t rank = scan[i] - 1;
    const uint32_t row = (uint32_t)(keys[i] >> 32);
    const uint32_t slot = rank + dir_before[row];
    if (head[i]) {
        rows[slot] = row;
        cols[slot] = (uint32_t)keys[i];
        if (i + 1 < n && !head[i + 1]) {   // key with more than one contribution: point at its run in the extras arrays
            const uint32_t k = i - rank;   // rank of element i+1 among the non-heads
            src1[slot] = 0x80000000u | k;
            extra_first[k] = srcs[i];
        } else src1[slot] = srcs[i];
    } else {
        const uint32_t k = i - rank - 1;
        extra_slot[k] = slot;
        extra_src[k] = srcs[i];
    }
}

// One CTA per LOCAL block: the rows of its direct DoFs.  Row of the k-th smallest DoF: columns = the DoFs k, k+1, ... of the sorted
// list, source = the pair's entry in the (canonically indexed, upper-triangular) value tile of the block's class.
__global__ void __launch_bounds__(256) emit_direct_kernel(const DevBlock* __restrict__ blocks, const uint32_t* __restrict__ canon_dof,
                                                          const uint32_t* __restrict__ cnt, const uint16_t* __restrict__ perm,
                                                          const uint32_t* __restrict__ row_ptr,
                                                          uint32_t* __restrict__ rows, uint32_t* __restrict__ cols, uint32_t* __restrict__ src1) {
    const DevBlock b = blocks[blockIdx.x];
    if (!b.local) return;
    __shared__ uint32_t s_dof[ORDER_MAX];
    __shared__ uint16_t s_can[ORDER_MAX];
    for (uint32_t k = threadIdx.x; k < b.nP; k += blockDim.x) { const uint16_t a = perm[b.dofP_off + k]; s_can[k] = a; s_dof[k] = canon_dof[b.dofP_off + a]; }
    __syncthreads();
    for (uint32_t k = 0; k < b.nP; k++) {
        const uint32_t r = s_dof[k];
        if (cnt[r] != 1u) continue;            // CTA-uniform
        const uint32_t base = row_ptr[r], a = s_can[k];
        for (uint32_t j = k + threadIdx.x; j < b.nP; j += blockDim.x) {
            const uint32_t q = s_can[j], lo = a < q ? a : q, hi = a < q ? q : a;
            const uint32_t slot = base + (j - k);
            rows[slot] = r; cols[slot] = s_dof[j];
            src1[slot] = (uint32_t)(b.v_off + (unsigned long long)lo * b.nQ + hi);
        }
    }
}

// ---- packed form of the source map -----------------------------------------------------------------------------------------------
// The scatter kernel streams 16 B of values per slot; a 4 B source index next to them is 20 % of its HBM traffic.  Consecutive
// slots read neighbouring entries of V (a row of the pattern walks along a row or a column of one class's value tile), so per
// chunk of SRC_CHUNK slots the sources are stored as 16-bit offsets from the chunk's smallest source: 2 B per slot + 4 B per
// chunk, both read in one step (no dependent load in front of the gather).  A chunk whose sources span more than 2^16 entries,
// or that holds a slot with more than one contribution (high bit of src1), keeps the plain 32-bit form (chunk_base = PLAIN).
__global__ void __launch_bounds__(256) pack_sources_kernel(const uint32_t* __restrict__ src1, unsigned long long nnz, unsigned long long n_chunks,
                                                          uint32_t* __restrict__ chunk_base, uint16_t* __restrict__ src16, uint32_t* __restrict__ n_plain) {
    static_assert(SRC_CHUNK == 64, "one warp packs one chunk, two slots per lane");
    const unsigned long long chunk = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) / 32;
    if (chunk >= n_chunks) return;
    const uint32_t lane = threadIdx.x % 32;
    const unsigned long long s0 = chunk * SRC_CHUNK + lane, s1 = s0 + 32;
    const bool in0 = s0 < nnz, in1 = s1 < nnz;
    const uint32_t v0 = in0 ? src1[s0] : 0u, v1 = in1 ? src1[s1] : 0u;
    uint32_t mn = min(in0 ? v0 : 0xffffffffu, in1 ? v1 : 0xffffffffu), mx = max(v0, v1);
    for (int d = 16; d; d >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d)); }
    const bool plain = (mx & 0x80000000u) || mx - mn > 0xffffu;
    if (lane == 0) { chunk_base[chunk] = plain ? SRC_CHUNK_PLAIN : mn; if (plain) atomicAdd(n_plain, 1u); }
    if (in0) src16[s0] = plain ? (uint16_t)0 : (uint16_t)(v0 - mn);
    if (in1) src16[s1] = plain ? (uint16_t)0 : (uint16_t)(v1 - mn);
}

// longest run of equal keys and number of keys with more than one contribution
__global__ void contrib_stats_kernel(const uint32_t* __restrict__ extra_slot, uint32_t n_extra, uint32_t* __restrict__ stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_extra) return;
    if (i == 0 || extra_slot[i - 1] != extra_slot[i]) {
        uint32_t run = 1;
        while (i + run < n_extra && extra_slot[i + run] == extra_slot[i]) run++;
        atomicMax(&stats[0], run + 1);
        atomicAdd(&stats[1], 1u);
    }
}

// nnz-balanced split of the slot interval [lo, hi) into `world` row-aligned blocks
__global__ void row_bounds_kernel(const uint32_t* __restrict__ rows, unsigned long long lo, unsigned long long hi, uint32_t world, unsigned long long* bounds) {
    const uint32_t r = threadIdx.x;
    if (r > world) return;
    if (r == 0) { bounds[0] = lo; return; }
    if (r == world) { bounds[world] = hi; return; }
    unsigned long long s = lo + (hi - lo) * r / world;
    while (s < hi && s > lo && rows[s] == rows[s - 1]) s++;   // advance to the next row start
    bounds[r] = s;
}
__global__ void first_slot_of_row_kernel(const uint32_t* __restrict__ rows, unsigned long long nnz, uint32_t row, unsigned long long* out) {
    unsigned long long lo = 0, hi = nnz;   // first slot with rows[slot] >= row
    while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (rows[mid] < row) lo = mid + 1; else hi = mid; }
    *out = lo;
}

}  // namespace

namespace {
struct DevBlk { void* p; size_t bytes; int dev; };
std::mutex g_dev_mu;
std::vector<DevBlk> g_dev_free;                       // cached free blocks
std::vector<DevBlk> g_dev_live;                       // sizes of the blocks handed out (looked up at dev_free)
size_t g_dev_cached = 0;
constexpr size_t DEV_CACHE_BYTES = (size_t)8 << 30;   // at most this much memory parked in the cache
constexpr size_t DEV_CACHE_BLOCKS = 96;
}  // namespace

namespace { extern cudaMemPool_t g_pool[64]; void pinned_trim(); }

cudaError_t dev_malloc(void** p, size_t bytes, cudaStream_t st) {
    bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        size_t best = SIZE_MAX;   // best fit among the blocks that waste at most 1/8 (+ 1 MB)
        for (size_t k = 0; k < g_dev_free.size(); k++) {
            const DevBlk& b = g_dev_free[k];
            if (b.dev == dev && b.bytes >= bytes && b.bytes <= bytes + bytes / 8 + (1u << 20) && (best == SIZE_MAX || b.bytes < g_dev_free[best].bytes)) best = k;
        }
        if (best != SIZE_MAX) {
            const DevBlk b = g_dev_free[best];
            g_dev_free.erase(g_dev_free.begin() + best);
            g_dev_cached -= b.bytes;
            g_dev_live.push_back(b);
            *p = b.p;
            return cudaSuccess;
        }
    }
    cudaMemPool_t pool = (dev >= 0 && dev < 64) ? g_pool[dev] : nullptr;
    const cudaError_t e = pool ? cudaMallocFromPoolAsync(p, bytes, pool, st) : cudaMallocAsync(p, bytes, st);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(g_dev_mu); g_dev_live.push_back(DevBlk{*p, bytes, dev}); }
    return e;
}

void dev_free(void* p, cudaStream_t st) {
    if (!p) return;
    DevBlk b{p, 0, 0};
    bool cache = false;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        for (size_t k = g_dev_live.size(); k-- > 0;)
            if (g_dev_live[k].p == p) { b = g_dev_live[k]; g_dev_live.erase(g_dev_live.begin() + k); break; }
        cache = b.bytes && g_dev_free.size() < DEV_CACHE_BLOCKS && g_dev_cached + b.bytes <= DEV_CACHE_BYTES;
        if (cache) { g_dev_free.push_back(b); g_dev_cached += b.bytes; }
    }
    if (!cache) cudaFreeAsync(p, st);
}

namespace {
cudaMemPool_t g_pool[64] = {};   // the library's own stream-ordered pool per device: the application's default pool keeps its settings
}

void dev_pool_init(int device) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (device < 0 || device >= 64 || g_pool[device]) return;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
        unsigned long long thr = ~0ull;   // keep freed memory in the pool: a one-shot caller builds and frees a plan per call
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        g_pool[device] = pool;
    } else cudaGetLastError();
}

// Returns the cached device blocks and pinned staging buffers to the driver (fem2d_trim_cache).
void dev_cache_trim() {
    std::vector<DevBlk> blocks;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        blocks.swap(g_dev_free);
        g_dev_cached = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (const DevBlk& b : blocks) { cudaSetDevice(b.dev); cudaFreeAsync(b.p, nullptr); }
    for (int d = 0; d < 64; d++)
        if (g_pool[d]) { cudaSetDevice(d); cudaStreamSynchronize(nullptr); cudaMemPoolTrimTo(g_pool[d], 0); }
    cudaSetDevice(cur);
    pinned_trim();
}

namespace {
// Small process-wide cache of pinned staging buffers: a one-shot call builds and frees a plan every time, and pinning a few MB
// costs about as much as the copy it serves.
std::mutex g_pin_mu;
std::vector<std::pair<void*, size_t>> g_pin_free;
void* pinned_acquire(size_t bytes, size_t* cap) {
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        size_t best = SIZE_MAX;   // best fit, so that the same-sized requests of successive plans land on the same buffers
        for (size_t k = 0; k < g_pin_free.size(); k++)
            if (g_pin_free[k].second >= bytes && (best == SIZE_MAX || g_pin_free[k].second < g_pin_free[best].second)) best = k;
        if (best != SIZE_MAX) { void* p = g_pin_free[best].first; *cap = g_pin_free[best].second; g_pin_free.erase(g_pin_free.begin() + best); return p; }
    }
    void* p = nullptr;
    bytes = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);   // MiB granularity: a slightly larger Domain re-uses the buffer
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *cap = bytes;
    return p;
}
void pinned_trim() {
    std::vector<std::pair<void*, size_t>> bufs;
    { std::lock_guard<std::mutex> lk(g_pin_mu); bufs.swap(g_pin_free); }
    for (auto& b : bufs) cudaFreeHost(b.first);
}
void pinned_release(void* p, size_t cap) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (g_pin_free.size() < 64) { g_pin_free.push_back({p, cap}); return; }   // (8 buffers per plan-building thread of a multi-device call)
    cudaFreeHost(p);
}
}  // namespace

namespace {
// Carves 256-byte aligned sub-buffers out of one allocation.
struct Arena {
    size_t size = 0;
    size_t reserve(size_t bytes) { const size_t off = size; size += (bytes + 255) & ~(size_t)255; return off; }
};
template <class T> T* at(void* base, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + off); }
}  // namespace

// Two allocations per symbolic call: the plan-owned arena (pattern, source map, descriptors) and one scratch arena (keys, sort
// double buffers, flags, scan, cub temp) that goes back to the stream-ordered pool at the end -- repeated calls on same-sized
// Domains re-use both blocks without touching the driver allocator.  All small host arrays travel in one staged H2D copy.
int device_symbolic(Plan& P, std::string& err) {
    const HostPlan& H = P.host;
    CK(cudaSetDevice(P.device));
    dev_pool_init(P.device);
    CK(cudaDeviceGetAttribute(&P.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, P.device));
    CK(cudaDeviceGetAttribute(&P.sm_count, cudaDevAttrMultiProcessorCount, P.device));
    P.split = split_items(H, H.items, H.packs);

    // ---- descriptor blob (plan-owned part first, then the scratch-only part), laid out identically on host and device
    std::vector<DevBlock> hb(H.blocks.size());
    for (size_t k = 0; k < H.blocks.size(); k++) {
        const BlockDesc& b = H.blocks[k];
        const ClassDesc& c = H.classes[b.cls];
        hb[k] = DevBlock{b.pair_off, c.v_off, H.bs_off[b.elemP], H.bs_off[b.elemQ], H.lists[c.listP].n, H.lists[c.listQ].n, c.local, 0};
    }
    Arena desc;
    const size_t o_classes = desc.reserve(H.classes.size() * sizeof(ClassDesc)), o_lists = desc.reserve(H.lists.size() * sizeof(ListDesc));
    const size_t o_si = desc.reserve(H.spec_i.size()), o_sj = desc.reserve(H.spec_j.size());
    const size_t o_tabs = desc.reserve(H.tables.size() * sizeof(TableDesc)), o_grams = desc.reserve(H.grams.size() * sizeof(GramDesc));
    const size_t o_items = desc.reserve(H.items.size() * sizeof(WorkItem));
    const size_t o_packs = desc.reserve(H.packs.size() * sizeof(PackDesc));
    const size_t o_glq = desc.reserve(4 * 128 * sizeof(double));
    const size_t o_ctr = desc.reserve(256);
    std::vector<uint32_t> h_voff(H.classes.size() + 1), h_mtoff(H.classes.size() + 1);
    {
        uint64_t mt = 0;
        for (size_t c = 0; c < H.classes.size(); c++) { h_voff[c] = (uint32_t)H.classes[c].v_off; h_mtoff[c] = (uint32_t)mt; mt += H.classes[c].n_mt; }
        h_voff[H.classes.size()] = (uint32_t)H.n_values; h_mtoff[H.classes.size()] = (uint32_t)mt;
        P.total_mt = mt;
        if (mt >= (1ull << 32)) { err = "too many micro-tiles"; return FEM2D_ERR_UNSUPPORTED; }
    }
    const size_t o_voff = desc.reserve(h_voff.size() * 4), o_mtoff = desc.reserve(h_mtoff.size() * 4);
    const size_t o_geom = desc.reserve(H.classes.size() * sizeof(ClassGeom));   // written on the device
    const size_t desc_plan_bytes = desc.size;                     // everything above lives as long as the plan
    const size_t o_blocks = desc.reserve(hb.size() * sizeof(DevBlock)), o_canon = desc.reserve(H.canon_dof.size() * 4);
    // staged in a cached pinned buffer: one asynchronous H2D at PCIe speed instead of a pageable copy
    size_t blob_cap = 0;
    unsigned char* blob = (unsigned char*)pinned_acquire(desc.size, &blob_cap);
    if (!blob) { err = "pinned host allocation failed"; return FEM2D_ERR_OUT_OF_MEMORY; }
    struct BlobGuard { unsigned char* p; size_t cap; ~BlobGuard() { pinned_release(p, cap); } } blob_guard{blob, blob_cap};   // released after the last sync below
    auto put = [&](size_t off, const void* src, size_t n) { if (n) std::memcpy(blob + off, src, n); };
    put(o_classes, H.classes.data(), H.classes.size() * sizeof(ClassDesc)); put(o_lists, H.lists.data(), H.lists.size() * sizeof(ListDesc));
    put(o_si, H.spec_i.data(), H.spec_i.size()); put(o_sj, H.spec_j.data(), H.spec_j.size());
    put(o_tabs, H.tables.data(), H.tables.size() * sizeof(TableDesc)); put(o_grams, H.grams.data(), H.grams.size() * sizeof(GramDesc));
    put(o_items, H.items.data(), H.items.size() * sizeof(WorkItem));
    put(o_packs, H.packs.data(), H.packs.size() * sizeof(PackDesc));
    put(o_voff, h_voff.data(), h_voff.size() * 4); put(o_mtoff, h_mtoff.data(), h_mtoff.size() * 4);
    put(o_blocks, hb.data(), hb.size() * sizeof(DevBlock)); put(o_canon, H.canon_dof.data(), H.canon_dof.size() * 4);

    // ---- scratch arena 1: per-DoF and per-block bookkeeping
    const uint32_t nd = H.n_dofs, nb = (uint32_t)hb.size();
    if (H.max_list_n > ORDER_MAX) { err = "BasisSpec list longer than 1024 functions"; return FEM2D_ERR_UNSUPPORTED; }
    size_t temp_small = 0, tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, temp_small, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(nd + 1)));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(nb + 1)));
    temp_small = std::max(temp_small, tb);
    Arena sc;
    const size_t s_desc = sc.reserve(desc.size - desc_plan_bytes);
    const size_t s_cnt = sc.reserve((size_t)nd * 4);                                           // blocks per DoF
    const size_t s_len = sc.reserve(((size_t)nd + 1) * 4), s_ldir = sc.reserve(((size_t)nd + 1) * 4);   // row lengths: all rows / direct rows only
    const size_t s_dbef = sc.reserve(((size_t)nd + 1) * 4);                                    // direct slots before each row
    const size_t s_perm = sc.reserve(H.canon_dof.size() * 2);                                  // sorted order of every Elem's DoFs
    const size_t s_bcnt = sc.reserve(((size_t)nb + 1) * 4), s_boff = sc.reserve(((size_t)nb + 1) * 4);
    const size_t s_tmp1 = sc.reserve(temp_small), s_stats = sc.reserve(8);
    void* scratch = nullptr;
    void* scratch2 = nullptr;
    CK(dev_malloc(&scratch, sc.size));
#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); dev_free(scratch); dev_free(scratch2); return FEM2D_ERR_CUDA; } } while (0)
    // descriptors: the plan-owned part goes to a first small plan allocation, the rest into the scratch arena
    CKC(dev_malloc(&P.d_desc_arena, desc_plan_bytes));
    CKC(cudaMemcpyAsync(P.d_desc_arena, blob, o_geom, cudaMemcpyHostToDevice, nullptr));   // the class constants behind o_geom are computed on the device
    CKC(cudaMemcpyAsync(at<char>(scratch, s_desc), blob + desc_plan_bytes, desc.size - desc_plan_bytes, cudaMemcpyHostToDevice, nullptr));
    P.d_classes = at<ClassDesc>(P.d_desc_arena, o_classes); P.d_lists = at<ListDesc>(P.d_desc_arena, o_lists);
    P.d_spec_i = at<uint8_t>(P.d_desc_arena, o_si); P.d_spec_j = at<uint8_t>(P.d_desc_arena, o_sj);
    P.d_tables = at<TableDesc>(P.d_desc_arena, o_tabs); P.d_grams = at<GramDesc>(P.d_desc_arena, o_grams);
    P.d_items = at<WorkItem>(P.d_desc_arena, o_items); P.d_glq = at<double>(P.d_desc_arena, o_glq);
    P.d_work_counter = at<uint32_t>(P.d_desc_arena, o_ctr);
    P.d_packs = at<PackDesc>(P.d_desc_arena, o_packs);
    P.d_class_voff = at<uint32_t>(P.d_desc_arena, o_voff); P.d_class_mtoff = at<uint32_t>(P.d_desc_arena, o_mtoff);
    P.d_class_geom = at<ClassGeom>(P.d_desc_arena, o_geom);
    CKC(launch_class_geom(P, (uint32_t)H.classes.size(), nullptr));
    DevBlock* d_blocks = at<DevBlock>(scratch, s_desc + (o_blocks - desc_plan_bytes));
    uint32_t* d_canon = at<uint32_t>(scratch, s_desc + (o_canon - desc_plan_bytes));
    uint32_t *d_cnt = at<uint32_t>(scratch, s_cnt), *d_len = at<uint32_t>(scratch, s_len), *d_ldir = at<uint32_t>(scratch, s_ldir);
    uint32_t *d_dbef = at<uint32_t>(scratch, s_dbef), *d_bcnt = at<uint32_t>(scratch, s_bcnt), *d_boff = at<uint32_t>(scratch, s_boff);
    uint32_t* d_stats = at<uint32_t>(scratch, s_stats);
    uint16_t* d_perm = at<uint16_t>(scratch, s_perm);
    void* d_tmp1 = at<char>(scratch, s_tmp1);

    // ---- which rows are direct, their lengths, and how many pairs go through the sort
    CKC(cudaMemsetAsync(at<char>(scratch, s_cnt), 0, s_dbef - s_cnt, nullptr));   // cnt, len, ldir
    CKC(cudaMemsetAsync(d_bcnt + nb, 0, 4, nullptr));
    incidence_kernel<<<nb, 128>>>(d_blocks, d_canon, d_cnt);
    CKC(cudaGetLastError());
    local_order_kernel<<<nb, 256>>>(d_blocks, d_canon, d_cnt, d_perm, d_len, d_ldir);
    CKC(cudaGetLastError());
    shared_pairs_kernel<false><<<nb, 256>>>(d_blocks, d_canon, d_cnt, d_bcnt, nullptr, nullptr, nullptr);
    CKC(cudaGetLastError());
    CKC(cub::DeviceScan::ExclusiveSum(d_tmp1, temp_small, d_bcnt, d_boff, (int)(nb + 1)));
    uint32_t n_sub = 0;
    CKC(cudaMemcpy(&n_sub, d_boff + nb, 4, cudaMemcpyDeviceToHost));

    // ---- scratch arena 2: the shared pairs, sorted by [row, col] (stable LSD radix sort on the significant key bits only)
    int bits = 1; while ((1ull << bits) < (unsigned long long)nd) bits++;
    size_t temp_bytes = 0, scan_bytes = 0;
    {
        cub::DoubleBuffer<unsigned long long> kq(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> vq(nullptr, nullptr);
        CKC(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, kq, vq, (int)std::max(n_sub, 1u), 0, 32 + bits));
        CKC(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)std::max(n_sub, 1u)));
        temp_bytes = std::max(temp_bytes, scan_bytes);
    }
    Arena s2;
    const size_t ns = std::max(n_sub, 1u);
    const size_t s_keys = s2.reserve(ns * 8), s_keys2 = s2.reserve(ns * 8), s_srcs = s2.reserve(ns * 4), s_srcs2 = s2.reserve(ns * 4);
    const size_t s_head = s2.reserve(ns * 4), s_scan = s2.reserve(ns * 4), s_temp = s2.reserve(temp_bytes);
    CKC(dev_malloc(&scratch2, s2.size));
    unsigned long long *d_keys = at<unsigned long long>(scratch2, s_keys), *d_keys2 = at<unsigned long long>(scratch2, s_keys2);
    uint32_t *d_srcs = at<uint32_t>(scratch2, s_srcs), *d_srcs2 = at<uint32_t>(scratch2, s_srcs2);
    uint32_t *d_head = at<uint32_t>(scratch2, s_head), *d_scan = at<uint32_t>(scratch2, s_scan);
    void* d_temp = at<char>(scratch2, s_temp);
    const unsigned long long* keys_sorted = d_keys;
    const uint32_t* srcs_sorted = d_srcs;
    uint32_t n_heads = 0;
    const unsigned gb = (n_sub + 255) / 256;
    if (n_sub) {
        shared_pairs_kernel<true><<<nb, 256>>>(d_blocks, d_canon, d_cnt, nullptr, d_boff, d_keys, d_srcs);
        CKC(cudaGetLastError());
        cub::DoubleBuffer<unsigned long long> kb(d_keys, d_keys2);
        cub::DoubleBuffer<uint32_t> vb(d_srcs, d_srcs2);
        CKC(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb, vb, (int)n_sub, 0, bits));
        CKC(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb, vb, (int)n_sub, 32, 32 + bits));
        keys_sorted = kb.Current(); srcs_sorted = vb.Current();
        head_flags_kernel<<<gb, 256>>>(keys_sorted, n_sub, d_head, d_len);
        CKC(cudaGetLastError());
        CKC(cub::DeviceScan::InclusiveSum(d_temp, temp_bytes, d_head, d_scan, (int)n_sub));
        CKC(cudaMemcpyAsync(&n_heads, d_scan + (n_sub - 1), 4, cudaMemcpyDeviceToHost, nullptr));
    }

    // ---- plan-owned pattern arena; the CSR row offsets are the scan of the row lengths
    Arena pr;
    const size_t p_rp = pr.reserve(((size_t)nd + 1) * 4);
    CKC(dev_malloc(&P.d_rowptr_arena, pr.size));
    P.d_row_ptr = at<uint32_t>(P.d_rowptr_arena, p_rp);
    CKC(cub::DeviceScan::ExclusiveSum(d_tmp1, temp_small, d_len, P.d_row_ptr, (int)(nd + 1)));
    CKC(cub::DeviceScan::ExclusiveSum(d_tmp1, temp_small, d_ldir, d_dbef, (int)(nd + 1)));
    uint32_t nnz32 = 0;
    CKC(cudaMemcpy(&nnz32, P.d_row_ptr + nd, 4, cudaMemcpyDeviceToHost));   // synchronises: n_heads has arrived too
    P.nnz = nnz32; P.nnz32_sentinel = nnz32; P.n_extra = (uint64_t)n_sub - n_heads;
    Arena pa;
    const size_t p_rows = pa.reserve((size_t)nnz32 * 4), p_cols = pa.reserve((size_t)nnz32 * 4), p_src1 = pa.reserve((size_t)nnz32 * 4);
    const size_t p_es = pa.reserve(P.n_extra * 4), p_ex = pa.reserve(P.n_extra * 4), p_ef = pa.reserve(P.n_extra * 4);
    CKC(dev_malloc(&P.d_pattern_arena, pa.size));
    P.d_rows = at<uint32_t>(P.d_pattern_arena, p_rows); P.d_cols = at<uint32_t>(P.d_pattern_arena, p_cols); P.d_src1 = at<uint32_t>(P.d_pattern_arena, p_src1);
    P.d_extra_slot = at<uint32_t>(P.d_pattern_arena, p_es); P.d_extra_src = at<uint32_t>(P.d_pattern_arena, p_ex); P.d_extra_first = at<uint32_t>(P.d_pattern_arena, p_ef);
    emit_direct_kernel<<<nb, 256>>>(d_blocks, d_canon, d_cnt, d_perm, P.d_row_ptr, P.d_rows, P.d_cols, P.d_src1);
    CKC(cudaGetLastError());
    if (n_sub) {
        emit_shared_kernel<<<gb, 256>>>(keys_sorted, srcs_sorted, d_head, d_scan, n_sub, d_dbef, P.d_rows, P.d_cols, P.d_src1, P.d_extra_slot, P.d_extra_src,
                                        P.d_extra_first);
        CKC(cudaGetLastError());
    }
    uint32_t stats[2] = {1, 0};
    CKC(cudaMemcpyAsync(d_stats, stats, 8, cudaMemcpyHostToDevice, nullptr));
    if (P.n_extra) {
        contrib_stats_kernel<<<(unsigned)((P.n_extra + 255) / 256), 256>>>(P.d_extra_slot, (uint32_t)P.n_extra, d_stats);
        CKC(cudaGetLastError());
    }
    CKC(cudaMemcpy(stats, d_stats, 8, cudaMemcpyDeviceToHost));   // synchronises the null stream: everything above is done
    P.max_contrib = stats[0]; P.n_multi = stats[1];
    P.n_sorted_pairs = n_sub;

    // ---- packed form of the source map (what the scatter kernel reads; src1 stays for the chunks that do not pack)
    {
        const unsigned long long n_chunks = ((unsigned long long)nnz32 + SRC_CHUNK - 1) / SRC_CHUNK;
        Arena ra;
        const size_t r_base = ra.reserve((n_chunks + 1) * 4), r_16 = ra.reserve(((size_t)nnz32 + 1) * 2);
        CKC(dev_malloc(&P.d_pack_arena, ra.size));
        P.d_chunk_base = at<uint32_t>(P.d_pack_arena, r_base); P.d_src16 = at<uint16_t>(P.d_pack_arena, r_16);
        if (nnz32) {
            CKC(cudaMemsetAsync(d_stats, 0, 4, nullptr));
            pack_sources_kernel<<<(unsigned)((n_chunks * 32 + 255) / 256), 256>>>(P.d_src1, nnz32, n_chunks, P.d_chunk_base, P.d_src16, d_stats);
            CKC(cudaGetLastError());
            uint32_t n_plain = 0;
            CKC(cudaMemcpy(&n_plain, d_stats, 4, cudaMemcpyDeviceToHost));
            P.n_plain_chunks = n_plain;
        }
    }
    dev_free(scratch); dev_free(scratch2);
#undef CKC
    return FEM2D_OK;
}

namespace {
// Marks the micro-tile that produces V[v]: inverse of the integrator's tile enumeration (plan_types.h make_subblocks).
__device__ void mark_source(uint32_t v, const ClassDesc* classes, const ListDesc* lists, const uint32_t* voff, const uint32_t* mtoff, uint32_t n_classes,
                            uint32_t tp, unsigned char* flags) {
    uint32_t lo = 0, hi = n_classes;                      // last class with voff <= v
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (voff[mid] <= v) lo = mid; else hi = mid; }
    const ClassDesc& c = classes[lo];
    const uint32_t nP = lists[c.listP].n, nUP = lists[c.listP].nU, nQ = lists[c.listQ].n, nUQ = lists[c.listQ].nU;
    const uint32_t t = v - voff[lo], a = t / nQ, b = t - a * nQ;
    const SubBlocks sb = make_subblocks(nP, nUP, nQ, nUQ, c.local, tp);
    const uint32_t idx = encode_tile(sb, a, b, nUP, nUQ, tp);
    flags[mtoff[lo] + idx] = 1;
}

__global__ void mark_tiles_kernel(const uint32_t* __restrict__ src1, const uint32_t* __restrict__ extra_slot, const uint32_t* __restrict__ extra_src,
                                  const uint32_t* __restrict__ extra_first, unsigned long long n_extra, unsigned long long begin, unsigned long long end,
                                  const ClassDesc* classes, const ListDesc* lists, const uint32_t* voff, const uint32_t* mtoff, uint32_t n_classes,
                                  uint32_t tp, unsigned char* flags) {
    const unsigned long long slot = begin + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (slot >= end) return;
    const uint32_t s = src1[slot];
    if (!(s & 0x80000000u)) { mark_source(s, classes, lists, voff, mtoff, n_classes, tp, flags); return; }
    unsigned long long k = s & 0x7fffffffu;
    mark_source(extra_first[k], classes, lists, voff, mtoff, n_classes, tp, flags);
    do { mark_source(extra_src[k], classes, lists, voff, mtoff, n_classes, tp, flags); k++; } while (k < n_extra && extra_slot[k] == slot);
}
}  // namespace

ItemSplit split_items(const HostPlan& H, const std::vector<WorkItem>& items, const std::vector<PackDesc>& packs) {
    // `items` is in launch order: packs (pack_items) or, without them, the big items as a prefix (order_items)
    ItemSplit sp;
    for (const WorkItem& it : items) {
        const uint32_t s = item_slab_stride(H, it);
        if (item_is_big(H, it)) { sp.n_big++; sp.stride_big = std::max(sp.stride_big, s); }
        else sp.stride_small = std::max(sp.stride_small, s);
    }
    sp.n_packs = (uint32_t)packs.size();
    for (const PackDesc& pk : packs) {
        uint32_t s = 0;
        for (uint32_t k = 0; k < pk.n; k++) s += item_slab_stride(H, items[pk.first + k]);
        sp.stride_pack = std::max(sp.stride_pack, s);
    }
    return sp;
}

int device_range_items(Plan& P, uint32_t n_ranges, const uint64_t* begins, const uint64_t* ends, const WorkItem** d_items, uint32_t* n_items,
                       const PackDesc** d_packs, ItemSplit* split, std::string& err) {
    if (n_ranges > MAX_SLOT_RANGES) { err = "too many slot ranges"; return FEM2D_ERR_BAD_ARGUMENT; }
    const bool full = n_ranges == 1 && begins[0] == 0 && ends[0] >= P.nnz;
    // restricting is pointless when the whole integrator is a single wave of CTAs anyway
    if (full || P.total_mt < (uint64_t)4 * 148 * K2_THREADS) { *d_items = P.d_items; *n_items = (uint32_t)P.host.items.size(); *d_packs = P.d_packs; *split = P.split; return FEM2D_OK; }
    bool same = P.d_range_items && n_ranges == P.range_n;
    for (uint32_t k = 0; same && k < n_ranges; k++) same = begins[k] == P.range_begin[k] && ends[k] == P.range_end[k];
    if (same) { *d_items = P.d_range_items; *n_items = P.n_range_items; *d_packs = P.d_range_packs; *split = P.range_split; return FEM2D_OK; }
    CK(cudaSetDevice(P.device));
    // the previous restricted list may still be read by an integrator launched on the caller's stream: nothing below is ordered
    // against that stream (null-stream kernels, blocking copies), so wait for the device before the list is replaced
    CK(cudaDeviceSynchronize());
    unsigned char* d_flags = nullptr;
    CK(dev_malloc((void**)&d_flags, P.total_mt));
    CK(cudaMemsetAsync(d_flags, 0, P.total_mt, nullptr));
    for (uint32_t k = 0; k < n_ranges; k++) {
        const uint64_t begin = begins[k], end = std::min<uint64_t>(ends[k], P.nnz);
        if (end <= begin) continue;
        mark_tiles_kernel<<<(unsigned)((end - begin + 255) / 256), 256>>>(P.d_src1, P.d_extra_slot, P.d_extra_src, P.d_extra_first, P.n_extra, begin, end,
                                                                           P.d_classes, P.d_lists, P.d_class_voff, P.d_class_mtoff,
                                                                           (uint32_t)P.host.classes.size(), P.host.tile_p, d_flags);
    }
    std::vector<unsigned char> flags(P.total_mt);
    cudaError_t e = cudaMemcpy(flags.data(), d_flags, P.total_mt, cudaMemcpyDeviceToHost);
    dev_free(d_flags);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return FEM2D_ERR_CUDA; }
    // items: per class, the runs of needed tiles (gaps of up to 2 tiles are bridged) packed into CTAs of <= ITEM_MAX_RANGES runs and
    // <= 2 * K2_THREADS tiles; every item records which function columns its tiles touch so it stages only those.
    std::vector<WorkItem> items;
    const uint32_t cap = K2_ROUNDS * (P.host.use_ws && P.host.tile_p == (uint32_t)K2_TILE_P ? P.host.ws_round_slots() / K2_WS_TPT : (uint32_t)K2_THREADS) - 31, gap = 2, tp = P.host.tile_p;   // tiles; + up to 31 slots of warp alignment (item_slots)
    uint64_t needed = 0, off = 0;
    std::vector<std::pair<uint32_t, uint32_t>> runs, cur;
    for (uint32_t c = 0; c < P.host.classes.size(); c++) {
        const ClassDesc& cd = P.host.classes[c];
        const uint32_t n = cd.n_mt;
        const unsigned char* f = flags.data() + off;
        off += n;
        runs.clear();
        for (uint32_t k = 0; k < n;) {
            if (!f[k]) { k++; continue; }
            uint32_t b = k, last = k;
            while (k < n && k - b < cap && (f[k] || k - last <= gap)) { if (f[k]) last = k; k++; }
            runs.push_back({b, last - b + 1});
            k = last + 1;
        }
        if (runs.empty()) continue;
        const ListDesc& LP = P.host.lists[cd.listP]; const ListDesc& LQ = P.host.lists[cd.listQ];
        const SubBlocks sb = make_subblocks(LP.n, LP.nU, LQ.n, LQ.nU, cd.local, tp);
        auto flush = [&]() {
            if (cur.empty()) return;
            uint32_t cols[2][2][2] = {{{UINT32_MAX, 0}, {UINT32_MAX, 0}}, {{UINT32_MAX, 0}, {UINT32_MAX, 0}}};
            for (auto& r : cur)
                for (uint32_t t = r.first; t < r.first + r.second; t++) {
                    uint32_t sub, rt, ct;
                    decode_tile(sb, t, tp, sub, rt, ct);
                    const uint32_t w = mt_width(sub);
                    const uint32_t r0 = rt * tp, r1 = std::min(r0 + tp, sb.rows[sub]), c0 = ct * w, c1 = std::min<uint32_t>(c0 + w, sb.cols[sub]);
                    uint32_t* pr = cols[0][sub >= 2]; uint32_t* qc = cols[1][sub & 1];
                    pr[0] = std::min(pr[0], r0); pr[1] = std::max(pr[1], r1);
                    qc[0] = std::min(qc[0], c0); qc[1] = std::max(qc[1], c1);
                }
            for (auto& sd : cols) for (auto& g : sd) if (g[0] == UINT32_MAX) g[0] = g[1] = 0;
            items.push_back(make_item(P.host, c, cur, cols));
            needed += items.back().mt_count;
            cur.clear();
        };
        uint32_t cnt = 0;
        for (auto& r : runs) {
            if (cur.size() == ITEM_MAX_RANGES || cnt + r.second > cap) { flush(); cnt = 0; }
            cur.push_back(r); cnt += r.second;
        }
        flush();
    }
    std::vector<PackDesc> packs;
    pack_items(P.host, items, packs);
    dev_free(P.d_range_items); P.d_range_items = nullptr; P.d_range_packs = nullptr;
    const size_t item_bytes = (std::max<size_t>(items.size(), 1) * sizeof(WorkItem) + 255) & ~(size_t)255;
    CK(dev_malloc((void**)&P.d_range_items, item_bytes + std::max<size_t>(packs.size(), 1) * sizeof(PackDesc)));
    P.d_range_packs = reinterpret_cast<PackDesc*>(reinterpret_cast<char*>(P.d_range_items) + item_bytes);
    CK(cudaMemcpy(P.d_range_items, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
    if (!packs.empty()) CK(cudaMemcpy(P.d_range_packs, packs.data(), packs.size() * sizeof(PackDesc), cudaMemcpyHostToDevice));
    P.n_range_items = (uint32_t)items.size(); P.range_n = n_ranges; P.range_mt_needed = needed;
    P.range_split = split_items(P.host, items, packs); *split = P.range_split;
    for (uint32_t k = 0; k < n_ranges; k++) { P.range_begin[k] = begins[k]; P.range_end[k] = ends[k]; }
    *d_packs = P.d_range_packs;
    *d_items = P.d_range_items; *n_items = P.n_range_items;
    return FEM2D_OK;
}

int device_first_slot_of_row(const Plan& P, uint32_t row, uint64_t* slot, std::string& err) {
    CK(cudaSetDevice(P.device));
    unsigned long long* d = nullptr;
    CK(dev_malloc((void**)&d, 8));
    first_slot_of_row_kernel<<<1, 1>>>(P.d_rows, P.nnz, row, d);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    dev_free(d);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return FEM2D_ERR_CUDA; }
    *slot = h;
    return FEM2D_OK;
}

int device_row_block_bounds(const Plan& P, uint32_t world, uint64_t* bounds, std::string& err) { return device_row_block_bounds_range(P, 0, P.nnz, world, bounds, err); }

int device_row_block_bounds_range(const Plan& P, uint64_t lo, uint64_t hi, uint32_t world, uint64_t* bounds, std::string& err) {
    if (world == 0 || world > 1023) { err = "world out of range"; return FEM2D_ERR_BAD_ARGUMENT; }
    CK(cudaSetDevice(P.device));
    unsigned long long* d_b = nullptr;
    CK(fem2d::dev_malloc((void**)&d_b, (world + 1) * 8));
    row_bounds_kernel<<<1, 1024>>>(P.d_rows, lo, hi, world, d_b);
    cudaError_t e = cudaMemcpy(bounds, d_b, (world + 1) * 8, cudaMemcpyDeviceToHost);
    fem2d::dev_free(d_b);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return FEM2D_ERR_CUDA; }
    return FEM2D_OK;
}

int device_row_ptr_host(Plan& P, cudaStream_t st, std::string& err) {
    if (P.h_row_ptr) return FEM2D_OK;
    CK(cudaSetDevice(P.device));
    P.h_row_ptr = (uint32_t*)pinned_acquire(((size_t)P.host.n_dofs + 1) * 4, &P.h_row_ptr_cap);
    if (!P.h_row_ptr) { err = "pinned host allocation failed"; return FEM2D_ERR_OUT_OF_MEMORY; }
    CK(cudaMemcpyAsync(P.h_row_ptr, P.d_row_ptr, ((size_t)P.host.n_dofs + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return FEM2D_OK;
}

namespace {
// slot s opens a run of consecutive column ids: first slot of its row, or its column is not the previous one + 1
struct ColRunHead {
    const uint32_t* rows; const uint32_t* cols;
    __device__ bool operator()(uint32_t s) const { return s == 0 || rows[s] != rows[s - 1] || cols[s] != cols[s - 1] + 1u; }
};
__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
}  // namespace

namespace {
// for every slot range [b, e): index of the last run that starts at or before b, and of the first run start >= e (the sentinel at the latest)
__global__ void run_window_kernel(const uint32_t* __restrict__ run_slot, uint32_t n_runs, const unsigned long long* __restrict__ be, uint32_t n_ranges,
                                  uint32_t* __restrict__ out) {
    const uint32_t k = threadIdx.x;
    if (k >= n_ranges) return;
    const unsigned long long b = be[2 * k], e = be[2 * k + 1];
    uint32_t lo = 0, hi = n_runs;                 // last index with run_slot[idx] <= b  (run_slot[0] == 0)
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (run_slot[mid] <= b) lo = mid; else hi = mid; }
    out[2 * k] = lo;
    uint32_t l2 = 0, h2 = n_runs;                 // first index with run_slot[idx] >= e  (run_slot[n_runs] == nnz)
    while (l2 < h2) { const uint32_t mid = (l2 + h2) >> 1; if (run_slot[mid] < e) l2 = mid + 1; else h2 = mid; }
    out[2 * k + 1] = l2;
}
}  // namespace

// Column runs of the pattern: compacted on the device once per plan (run starts + first columns stay on the device); the pinned host
// copies hold either everything (n_ranges == 0) or only the windows of runs that cover the given slot ranges -- a rank of a
// multi-device call expands 1/N of the pattern and fetches 1/N of the runs.  windows[2k], windows[2k+1]: first run and end run (its
// start is >= the range's end) of range k.
int device_col_runs_host(Plan& P, cudaStream_t st, uint32_t n_ranges, const uint64_t* begins, const uint64_t* ends, uint32_t* windows, std::string& err) {
    CK(cudaSetDevice(P.device));
    const uint32_t nnz = (uint32_t)P.nnz;
    if (!P.d_col_run_slot) {
        uint32_t *d_slot = nullptr, *d_col = nullptr;
        CK(dev_malloc((void**)&d_slot, ((size_t)nnz + 1) * 4 + 256, st));
        uint32_t* d_n = d_slot + (((size_t)nnz + 1 + 63) & ~(size_t)63);   // the count lives in the slack behind the slots
        cub::CountingInputIterator<uint32_t> it(0);
        ColRunHead pred{P.d_rows, P.d_cols};
        size_t temp = 0;
        CK(cub::DeviceSelect::If(nullptr, temp, it, d_slot, d_n, (int)nnz, pred, st));
        void* d_temp = nullptr;
        { const cudaError_t e0 = dev_malloc(&d_temp, temp, st); if (e0 != cudaSuccess) { dev_free(d_slot, st); err = cudaGetErrorString(e0); return FEM2D_ERR_CUDA; } }
        cudaError_t e = cub::DeviceSelect::If(d_temp, temp, it, d_slot, d_n, (int)nnz, pred, st);
        uint32_t n_runs = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&n_runs, d_n, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = dev_malloc((void**)&d_col, ((size_t)n_runs + 1) * 4, st);
        if (e == cudaSuccess && n_runs) { gather_u32_kernel<<<(n_runs + 255) / 256, 256, 0, st>>>(P.d_cols, d_slot, n_runs, d_col); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_slot + n_runs, &P.nnz32_sentinel, 4, cudaMemcpyHostToDevice, st);   // sentinel: end of the last run
        if (e == cudaSuccess) {
            P.h_col_run_slot = (uint32_t*)pinned_acquire(((size_t)n_runs + 1) * 4, &P.h_col_run_cap[0]);
            P.h_col_run_col = (uint32_t*)pinned_acquire(((size_t)n_runs + 1) * 4, &P.h_col_run_cap[1]);
            if (!P.h_col_run_slot || !P.h_col_run_col) e = cudaErrorMemoryAllocation;
        }
        dev_free(d_temp, st);
        if (e != cudaSuccess) {
            dev_free(d_slot, st); dev_free(d_col, st);
            pinned_release(P.h_col_run_slot, P.h_col_run_cap[0]); pinned_release(P.h_col_run_col, P.h_col_run_cap[1]);
            P.h_col_run_slot = P.h_col_run_col = nullptr;
            err = cudaGetErrorString(e); return FEM2D_ERR_CUDA;
        }
        P.d_col_run_slot = d_slot; P.d_col_run_col = d_col; P.n_col_runs = n_runs; P.col_runs_all = false;
    }
    const uint32_t n_runs = (uint32_t)P.n_col_runs;
    if (n_ranges == 0) {
        if (!P.col_runs_all) {
            CK(cudaMemcpyAsync(P.h_col_run_slot, P.d_col_run_slot, ((size_t)n_runs + 1) * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(P.h_col_run_col, P.d_col_run_col, (size_t)n_runs * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            P.col_runs_all = true;
        }
        return FEM2D_OK;
    }
    if (n_ranges > MAX_SLOT_RANGES) { err = "too many slot ranges"; return FEM2D_ERR_BAD_ARGUMENT; }
    if (P.col_runs_all) {   // everything is on the host already: the windows follow from the host copy
        for (uint32_t k = 0; k < n_ranges; k++) {
            const uint32_t* s0 = P.h_col_run_slot;
            windows[2 * k] = (uint32_t)(std::upper_bound(s0, s0 + n_runs + 1, (uint32_t)std::min<uint64_t>(begins[k], nnz)) - s0) - 1;
            windows[2 * k + 1] = (uint32_t)(std::lower_bound(s0, s0 + n_runs + 1, (uint32_t)std::min<uint64_t>(ends[k], nnz)) - s0);
        }
        return FEM2D_OK;
    }
    unsigned long long h_be[2 * MAX_SLOT_RANGES];
    for (uint32_t k = 0; k < n_ranges; k++) { h_be[2 * k] = begins[k]; h_be[2 * k + 1] = std::min<uint64_t>(ends[k], nnz); }
    unsigned long long* d_be = nullptr;
    CK(dev_malloc((void**)&d_be, sizeof(h_be) + 2 * MAX_SLOT_RANGES * 4, st));
    uint32_t* d_w = reinterpret_cast<uint32_t*>(d_be + 2 * MAX_SLOT_RANGES);
    cudaError_t e = cudaMemcpyAsync(d_be, h_be, sizeof(unsigned long long) * 2 * n_ranges, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) { run_window_kernel<<<1, 32, 0, st>>>(P.d_col_run_slot, n_runs, d_be, n_ranges, d_w); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(windows, d_w, 2 * n_ranges * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(d_be, st);
    for (uint32_t k = 0; k < n_ranges && e == cudaSuccess; k++) {
        const uint32_t lo = windows[2 * k], hi = windows[2 * k + 1];
        if (hi < lo) continue;
        e = cudaMemcpyAsync(P.h_col_run_slot + lo, P.d_col_run_slot + lo, ((size_t)hi - lo + 1) * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && hi > lo) e = cudaMemcpyAsync(P.h_col_run_col + lo, P.d_col_run_col + lo, ((size_t)hi - lo) * 4, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return FEM2D_ERR_CUDA; }
    return FEM2D_OK;
}

void device_plan_release(Plan& P) {
    if (P.device < 0) return;
    cudaSetDevice(P.device);
    cudaDeviceSynchronize();   // numeric work may still be in flight on a caller stream
    dev_free(P.d_desc_arena); dev_free(P.d_pattern_arena); dev_free(P.d_rowptr_arena);   // descriptors, GLQ buffer, pattern, source map
    dev_free(P.d_pack_arena);
    pinned_release(P.h_row_ptr, P.h_row_ptr_cap); P.h_row_ptr = nullptr;
    pinned_release(P.h_col_run_slot, P.h_col_run_cap[0]); pinned_release(P.h_col_run_col, P.h_col_run_cap[1]);
    P.h_col_run_slot = P.h_col_run_col = nullptr;
    dev_free(P.d_col_run_slot); dev_free(P.d_col_run_col);
    dev_free(P.d_range_items); dev_free(P.d_aij_arena);
    dev_free(P.d_V); dev_free(P.d_tabs); dev_free(P.d_gram); dev_free(P.d_dmma_items); dev_free(P.d_out_a); dev_free(P.d_out_b);
    for (int r = 0; r < Plan::RING; r++) for (int k = 0; k < 4; k++) if (P.ev[r][k]) cudaEventDestroy(P.ev[r][k]);
}

}  // namespace fem2d
