// Device part of the symbolic phase: pair keys -> radix sort -> unique = the fixed upper-triangular pattern
// (BTreeMap<[u32;2]> order, sparse_matrix.rs:16,48-58) and the per-slot source map that replaces the reference's
// per-Elem BTreeMap inserts and serial merge (sparse_matrix.rs:68-120, linalg.rs:59-81).
#include <cub/cub.cuh>

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

// One CTA per block of pairs: key = [min(dof_p, dof_q), max] (sparse_matrix.rs:78-91), src = index of the pair's value in V.
__global__ void keygen_kernel(const DevBlock* __restrict__ blocks, const uint32_t* __restrict__ canon_dof,
                              unsigned long long* __restrict__ keys, uint32_t* __restrict__ srcs) {
    const DevBlock b = blocks[blockIdx.x];
    const uint32_t* dp = canon_dof + b.dofP_off;
    const uint32_t* dq = canon_dof + b.dofQ_off;
    const uint32_t total = b.nP * b.nQ;
    for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
        const uint32_t a = t / b.nQ, q = t - a * b.nQ;
        unsigned long long pos;
        if (b.local) {
            if (q < a) continue;
            pos = b.pair_off + (unsigned long long)a * b.nQ - (unsigned long long)a * (a - 1) / 2 - a + q;   // packed upper triangle
        } else pos = b.pair_off + t;
        const uint32_t x = dp[a], y = dq[q];
        const uint32_t r = x < y ? x : y, c = x < y ? y : x;
        keys[pos] = (unsigned long long)r << 32 | c;
        srcs[pos] = (uint32_t)(b.v_off + t);
    }
}

__global__ void head_flags_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// slot_of[i] = (inclusive scan of head)[i] - 1.  Heads write the pattern, the others go to the extras list (kept in
// sorted order: position i - slot_of[i] - 1 is the rank among non-heads).
__global__ void emit_pattern_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ srcs,
                                    const uint32_t* __restrict__ head, const uint32_t* __restrict__ scan, uint32_t n,
                                    uint32_t* __restrict__ rows, uint32_t* __restrict__ cols, uint32_t* __restrict__ src1,
                                    uint32_t* __restrict__ extra_slot, uint32_t* __restrict__ extra_src, uint32_t* __restrict__ extra_first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = scan[i] - 1;
    if (head[i]) {
        rows[slot] = (uint32_t)(keys[i] >> 32);
        cols[slot] = (uint32_t)keys[i];
        if (i + 1 < n && !head[i + 1]) {   // key with more than one contribution: point at its run in the extras arrays
            const uint32_t k = i - slot;   // rank of element i+1 among the non-heads
            src1[slot] = 0x80000000u | k;
            extra_first[k] = srcs[i];
        } else src1[slot] = srcs[i];
    } else {
        const uint32_t k = i - slot - 1;
        extra_slot[k] = slot;
        extra_src[k] = srcs[i];
    }
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

__global__ void row_bounds_kernel(const uint32_t* __restrict__ rows, unsigned long long nnz, uint32_t world, unsigned long long* bounds) {
    const uint32_t r = threadIdx.x;
    if (r > world) return;
    if (r == 0) { bounds[0] = 0; return; }
    if (r == world) { bounds[world] = nnz; return; }
    unsigned long long s = nnz * r / world;
    while (s < nnz && s > 0 && rows[s] == rows[s - 1]) s++;   // advance to the next row start
    bounds[r] = s;
}

template <class T>
int upload(T*& dst, const T* src, size_t n, std::string& err) {
    dst = nullptr;
    if (n == 0) n = 1;
    CK(fem2d::dev_malloc((void**)&dst, n * sizeof(T)));
    if (src) CK(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return FEM2D_OK;
}

}  // namespace

void dev_pool_init(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device] = true;
}

int device_symbolic(Plan& P, std::string& err) {
    const HostPlan& H = P.host;
    CK(cudaSetDevice(P.device));
    dev_pool_init(P.device);
    CK(cudaDeviceGetAttribute(&P.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, P.device));
    CK(cudaDeviceGetAttribute(&P.sm_count, cudaDevAttrMultiProcessorCount, P.device));
    int st;
    if ((st = upload(P.d_classes, H.classes.data(), H.classes.size(), err))) return st;
    if ((st = upload(P.d_lists, H.lists.data(), H.lists.size(), err))) return st;
    if ((st = upload(P.d_spec_i, H.spec_i.data(), H.spec_i.size(), err))) return st;
    if ((st = upload(P.d_spec_j, H.spec_j.data(), H.spec_j.size(), err))) return st;
    if ((st = upload(P.d_tables, H.tables.data(), H.tables.size(), err))) return st;
    if ((st = upload(P.d_grams, H.grams.data(), H.grams.size(), err))) return st;
    if ((st = upload(P.d_items, H.items.data(), H.items.size(), err))) return st;

    const uint32_t np = (uint32_t)H.n_pairs;
    std::vector<DevBlock> hb(H.blocks.size());
    for (size_t k = 0; k < H.blocks.size(); k++) {
        const BlockDesc& b = H.blocks[k];
        const ClassDesc& c = H.classes[b.cls];
        hb[k] = DevBlock{b.pair_off, c.v_off, H.bs_off[b.elemP], H.bs_off[b.elemQ], H.lists[c.listP].n, H.lists[c.listQ].n, c.local, 0};
    }
    DevBlock* d_blocks = nullptr; uint32_t* d_canon = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_srcs = nullptr, *d_srcs2 = nullptr, *d_head = nullptr, *d_scan = nullptr, *d_stats = nullptr;
    void* d_temp = nullptr;
    auto cleanup = [&]() {
        fem2d::dev_free(d_blocks); fem2d::dev_free(d_canon); fem2d::dev_free(d_keys); fem2d::dev_free(d_keys2); fem2d::dev_free(d_srcs); fem2d::dev_free(d_srcs2);
        fem2d::dev_free(d_head); fem2d::dev_free(d_scan); fem2d::dev_free(d_stats); fem2d::dev_free(d_temp);
    };
#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); cleanup(); return FEM2D_ERR_CUDA; } } while (0)
    if ((st = upload(d_blocks, hb.data(), hb.size(), err))) { cleanup(); return st; }
    if ((st = upload(d_canon, H.canon_dof.data(), H.canon_dof.size(), err))) { cleanup(); return st; }
    CKC(fem2d::dev_malloc((void**)&d_keys, (size_t)np * 8)); CKC(fem2d::dev_malloc((void**)&d_keys2, (size_t)np * 8));
    CKC(fem2d::dev_malloc((void**)&d_srcs, (size_t)np * 4)); CKC(fem2d::dev_malloc((void**)&d_srcs2, (size_t)np * 4));
    keygen_kernel<<<(unsigned)hb.size(), 256>>>(d_blocks, d_canon, d_keys, d_srcs);
    CKC(cudaGetLastError());

    // stable LSD radix sort on the significant key bits only: [row | col] with bits(n_dofs) each
    int bits = 1; while ((1ull << bits) < (unsigned long long)H.n_dofs) bits++;
    cub::DoubleBuffer<unsigned long long> kb(d_keys, d_keys2);
    cub::DoubleBuffer<uint32_t> vb(d_srcs, d_srcs2);
    size_t temp_bytes = 0;
    CKC(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, kb, vb, (int)np, 0, 32 + bits));
    size_t scan_bytes = 0;
    CKC(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, d_head, d_scan, (int)np));
    temp_bytes = std::max(temp_bytes, scan_bytes);
    CKC(fem2d::dev_malloc(&d_temp, temp_bytes ? temp_bytes : 1));
    // keys use [row << 32 | col]: sort low 'bits' of col, then low 'bits' of row.  Two passes keep the bit count minimal.
    CKC(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb, vb, (int)np, 0, bits));
    CKC(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb, vb, (int)np, 32, 32 + bits));
    const unsigned long long* keys_sorted = kb.Current();
    const uint32_t* srcs_sorted = vb.Current();

    CKC(fem2d::dev_malloc((void**)&d_head, (size_t)np * 4)); CKC(fem2d::dev_malloc((void**)&d_scan, (size_t)np * 4));
    const unsigned gb = (np + 255) / 256;
    head_flags_kernel<<<gb, 256>>>(keys_sorted, np, d_head);
    CKC(cudaGetLastError());
    CKC(cub::DeviceScan::InclusiveSum(d_temp, temp_bytes, d_head, d_scan, (int)np));
    uint32_t nnz32 = 0;
    CKC(cudaMemcpy(&nnz32, d_scan + (np - 1), 4, cudaMemcpyDeviceToHost));
    P.nnz = nnz32; P.n_extra = (uint64_t)np - nnz32;
    CKC(fem2d::dev_malloc((void**)&P.d_rows, (size_t)nnz32 * 4)); CKC(fem2d::dev_malloc((void**)&P.d_cols, (size_t)nnz32 * 4));
    CKC(fem2d::dev_malloc((void**)&P.d_src1, (size_t)nnz32 * 4));
    CKC(fem2d::dev_malloc((void**)&P.d_extra_slot, (P.n_extra ? P.n_extra : 1) * 4)); CKC(fem2d::dev_malloc((void**)&P.d_extra_src, (P.n_extra ? P.n_extra : 1) * 4));
    CKC(fem2d::dev_malloc((void**)&P.d_extra_first, (P.n_extra ? P.n_extra : 1) * 4));
    emit_pattern_kernel<<<gb, 256>>>(keys_sorted, srcs_sorted, d_head, d_scan, np, P.d_rows, P.d_cols, P.d_src1, P.d_extra_slot, P.d_extra_src, P.d_extra_first);
    CKC(cudaGetLastError());
    uint32_t stats[2] = {1, 0};
    CKC(fem2d::dev_malloc((void**)&d_stats, 8));
    CKC(cudaMemcpy(d_stats, stats, 8, cudaMemcpyHostToDevice));
    if (P.n_extra) {
        contrib_stats_kernel<<<(unsigned)((P.n_extra + 255) / 256), 256>>>(P.d_extra_slot, (uint32_t)P.n_extra, d_stats);
        CKC(cudaGetLastError());
    }
    CKC(cudaMemcpy(stats, d_stats, 8, cudaMemcpyDeviceToHost));
    P.max_contrib = stats[0]; P.n_multi = stats[1];

    CKC(cudaDeviceSynchronize());
    cleanup();
#undef CKC
    for (int r = 0; r < Plan::RING; r++) for (int k = 0; k < 4; k++) CK(cudaEventCreate(&P.ev[r][k]));
    CK(fem2d::dev_malloc((void**)&P.d_glq, 4 * 128 * sizeof(double)));
    return FEM2D_OK;
}

int device_row_block_bounds(const Plan& P, uint32_t world, uint64_t* bounds, std::string& err) {
    if (world == 0 || world > 1023) { err = "world out of range"; return FEM2D_ERR_BAD_ARGUMENT; }
    CK(cudaSetDevice(P.device));
    unsigned long long* d_b = nullptr;
    CK(fem2d::dev_malloc((void**)&d_b, (world + 1) * 8));
    row_bounds_kernel<<<1, 1024>>>(P.d_rows, P.nnz, world, d_b);
    cudaError_t e = cudaMemcpy(bounds, d_b, (world + 1) * 8, cudaMemcpyDeviceToHost);
    fem2d::dev_free(d_b);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return FEM2D_ERR_CUDA; }
    return FEM2D_OK;
}

void device_plan_release(Plan& P) {
    if (P.device < 0) return;
    cudaSetDevice(P.device);
    cudaDeviceSynchronize();   // numeric work may still be in flight on a caller stream
    fem2d::dev_free(P.d_classes); fem2d::dev_free(P.d_lists); fem2d::dev_free(P.d_spec_i); fem2d::dev_free(P.d_spec_j); fem2d::dev_free(P.d_tables); fem2d::dev_free(P.d_grams); fem2d::dev_free(P.d_items);
    fem2d::dev_free(P.d_rows); fem2d::dev_free(P.d_cols); fem2d::dev_free(P.d_src1); fem2d::dev_free(P.d_extra_slot); fem2d::dev_free(P.d_extra_src); fem2d::dev_free(P.d_extra_first);
    fem2d::dev_free(P.d_V); fem2d::dev_free(P.d_tabs); fem2d::dev_free(P.d_glq); fem2d::dev_free(P.d_gram); fem2d::dev_free(P.d_dmma_items); fem2d::dev_free(P.d_out_a); fem2d::dev_free(P.d_out_b);
    for (int r = 0; r < Plan::RING; r++) for (int k = 0; k < 4; k++) if (P.ev[r][k]) cudaEventDestroy(P.ev[r][k]);
}

}  // namespace fem2d
