// Internal: the plan object behind `fem2d_plan` and the launch wrappers implemented in the .cu files.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "plan_host.hpp"

namespace fem2d {

struct ItemSplit { uint32_t n_big = 0, stride_big = 0, stride_small = 0, n_packs = 0, stride_pack = 0; };   // stride_pack: widest slab row (all segments) of a pack

struct Plan {
    int device = -1;   // -1: host-only plan (pattern + partitioning available, numeric entry points refuse)
    HostPlan host;
    HostPattern host_pattern;   // filled for host-only plans only
    uint64_t nnz = 0, n_extra = 0;
    uint32_t max_contrib = 0;
    uint64_t n_multi = 0;
    uint64_t t_host_us = 0, t_device_us = 0;   // wall time of the two halves of the symbolic phase

    // device-resident plan data (sub-buffers of the two arenas below)
    void* d_desc_arena = nullptr;
    void* d_pattern_arena = nullptr;
    void* d_rowptr_arena = nullptr;
    uint64_t n_sorted_pairs = 0;         // pairs whose row DoF is shared between blocks (the part of the pattern that is sorted)
    ClassDesc* d_classes = nullptr;
    ListDesc* d_lists = nullptr;
    uint8_t* d_spec_i = nullptr;
    uint8_t* d_spec_j = nullptr;
    TableDesc* d_tables = nullptr;
    GramDesc* d_grams = nullptr;
    WorkItem* d_items = nullptr;
    ClassGeom* d_class_geom = nullptr;   // desc arena; filled by launch_class_geom during the symbolic phase
    uint32_t* d_rows = nullptr;
    uint32_t* d_cols = nullptr;
    uint32_t* d_src1 = nullptr;
    uint32_t* d_row_ptr = nullptr;       // [n_dofs + 1] CSR row offsets of the pattern (first slot of every row)
    uint32_t* h_row_ptr = nullptr;       // pinned host copy, fetched on first use (host `rows` output is expanded from it)
    size_t h_row_ptr_cap = 0;
    // runs of consecutive column ids inside a row (built on first use by a host-output call): first slot and first column of each run,
    // pinned host copies; h_col_run_slot has n_col_runs + 1 entries (last = nnz)
    uint32_t* h_col_run_slot = nullptr; uint32_t* h_col_run_col = nullptr;
    size_t h_col_run_cap[2] = {0, 0};
    uint64_t n_col_runs = 0;
    uint32_t* d_col_run_slot = nullptr; uint32_t* d_col_run_col = nullptr;   // device copies (kept: later calls may need other windows)
    bool col_runs_all = false;           // the pinned host copies hold every run (otherwise only the windows fetched so far)
    uint32_t nnz32_sentinel = 0;         // == nnz, kept in the plan so an async H2D of it has a stable source
    uint32_t* d_extra_slot = nullptr;
    uint32_t* d_extra_src = nullptr;
    uint32_t* d_extra_first = nullptr;   // first contribution of a multi-contribution slot, stored at the head of its extras run
    // packed form of src1 (device_plan.cu pack_sources_kernel): 16-bit offsets from a per-chunk base; chunk_base == SRC_CHUNK_PLAIN
    // marks a chunk that keeps the 32-bit form
    void* d_pack_arena = nullptr;
    uint32_t* d_chunk_base = nullptr;
    uint16_t* d_src16 = nullptr;
    uint64_t n_plain_chunks = 0;

    // numeric scratch (allocated lazily, reused across calls)
    double2* d_V = nullptr;
    double* d_tabs = nullptr;
    size_t tabs_capacity = 0;   // doubles
    double* d_glq = nullptr;    // u_pts[128] u_w[128] v_pts[128] v_w[128]
    double h_glq[512] = {};     // host copy of what d_glq holds (skips the upload when the caller passes the same nodes again)
    bool glq_valid = false;
    double* d_gram = nullptr;   // fast modes scratch
    uint32_t* d_work_counter = nullptr;  // work-item counter of the persistent integrator (desc arena; reset by the sampler kernel)
    uint32_t* d_class_voff = nullptr;    // [n_classes + 1] first V entry of each class (desc arena)
    uint32_t* d_class_mtoff = nullptr;   // [n_classes + 1] first micro-tile of each class in the plan-wide tile numbering
    uint64_t total_mt = 0;
    // row-block restricted item list (multi-GPU sharding of the integrator), valid for [range_begin, range_end)
    uint32_t range_n = 0;
    uint64_t range_begin[4] = {0, 0, 0, 0}, range_end[4] = {0, 0, 0, 0};
    PackDesc* d_packs = nullptr;         // packs of host.items (desc arena)
    PackDesc* d_range_packs = nullptr;   // packs of the restricted item list (same allocation as d_range_items)
    WorkItem* d_range_items = nullptr;
    uint32_t n_range_items = 0;
    ItemSplit split, range_split;        // size split of host.items / of the restricted item list
    uint64_t range_mt_needed = 0;
    void* d_dmma_items = nullptr;   // tile work items of the DMMA integrator (built on first use)
    uint32_t n_dmma_items = 0;
    size_t gram_capacity = 0;
    // PETSc AIJ emission (petsc.cu): transpose of the strictly upper pattern and full-row offsets, built on first use
    void* d_aij_arena = nullptr;
    uint32_t *d_aij_lower_cnt = nullptr, *d_aij_lower_start = nullptr, *d_aij_counts = nullptr, *d_aij_full_start = nullptr, *d_aij_lower_slot = nullptr;
    uint32_t aij_n_lower = 0;
    uint64_t aij_nnz_full = 0;
    double* d_out_a = nullptr;  // staging for host-output calls
    double* d_out_b = nullptr;

    static constexpr int RING = 64;      // event sets of the last RING numeric calls (read back without syncing in between)
    cudaEvent_t ev[RING][4] = {};
    uint64_t n_calls = 0, n_timed_calls = 0;
    bool phase_timing = false;           // record the per-phase events (costs the launch overlap between the kernels)
    uint32_t last_launches[4] = {0, 0, 0, 0};
    int max_smem_optin = 0;
    int sm_count = 0;
};

// Device memory: blocks come from the stream-ordered pool of the device (cudaMallocAsync, unlimited release threshold) through a
// small size-matched cache of freed blocks.  A one-shot caller builds and frees a plan per call; the pool alone does not hand the
// same blocks back for the same sequence of requests (it splits large free blocks for small requests and then has to map fresh
// memory for the large ones: milliseconds, and erratic), the cache does.  Invariant: dev_free is only called once the block's last
// use has completed (after a synchronising call), so a cached block can be handed to any stream.
cudaError_t dev_malloc(void** p, size_t bytes, cudaStream_t st = nullptr);
void dev_free(void* p, cudaStream_t st = nullptr);
void dev_pool_init(int device);
void dev_cache_trim();   // give every cached device block and pinned staging buffer back to the driver

constexpr uint32_t MAX_SLOT_RANGES = 4;   // slot ranges one numeric call can cover (multi-GPU: a rank's Elem-type rows + its edge-type rows)
constexpr uint32_t SRC_CHUNK = 64;                     // slots per chunk of the packed source map
constexpr uint32_t SRC_CHUNK_PLAIN = 0x80000000u;      // chunk_base value of a chunk that is read through the plain 32-bit src1
constexpr uint32_t K3_THREADS = 128;          // scatter kernel: one CTA per K3_BLOCK_SLOTS-aligned block of slots,
constexpr uint32_t K3_ITERS = 4;              // K3_ITERS slots per thread
constexpr uint32_t K3_BLOCK_SLOTS = K3_THREADS * K3_ITERS;
constexpr uint32_t K3_PREFETCH_SLOTS = 1u << 20;   // distance of the L2 prefetch of the 16-bit offset stream
constexpr uint32_t MAX_GLQ = 128;   // default_ngq(20) = 128 (basis.rs:172-177)

// device_plan.cu
int device_symbolic(Plan& plan, std::string& err);
int device_row_block_bounds(const Plan& plan, uint32_t world, uint64_t* bounds, std::string& err);
// Work items of the exact integrator restricted to the micro-tiles that the slots [begin, end) read (cached per plan for the last
// range; the full item list is returned for the full range and for plans too small to be worth restricting).
int device_range_items(Plan& plan, uint32_t n_ranges, const uint64_t* begins, const uint64_t* ends, const WorkItem** d_items, uint32_t* n_items,
                       const PackDesc** d_packs, ItemSplit* split, std::string& err);
// First slot whose row is >= `row` (binary search over the device pattern).
int device_first_slot_of_row(const Plan& plan, uint32_t row, uint64_t* slot, std::string& err);
int device_row_block_bounds_range(const Plan& plan, uint64_t lo, uint64_t hi, uint32_t world, uint64_t* bounds, std::string& err);
// Pinned host copy of the CSR row offsets (fetched once per plan).
int device_row_ptr_host(Plan& plan, cudaStream_t st, std::string& err);
// Column runs of the pattern, compacted on the device and copied to pinned host memory (once per plan).
int device_col_runs_host(Plan& plan, cudaStream_t st, uint32_t n_ranges, const uint64_t* begins, const uint64_t* ends, uint32_t* windows, std::string& err);
void device_plan_release(Plan& plan);

// kernels_exact.cu  (compiled with -fmad=false)
cudaError_t launch_class_geom(const Plan& plan, uint32_t n_classes, cudaStream_t st);   // fills plan.d_class_geom from plan.d_classes
cudaError_t launch_k1_tables(const Plan& plan, int basis_kind, uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st);
cudaError_t launch_k2_exact(const Plan& plan, const WorkItem* d_items, uint32_t n_items, const PackDesc* d_packs, const ItemSplit& split,
                            uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st, uint32_t* launches);
// Size split of an item list ordered by mt_count (largest first): number of items that need K2_THREADS-wide CTAs and the widest
// slab row (pad4(U) + pad4(V) functions of P, + of Q unless local) among the classes of each part.
ItemSplit split_items(const HostPlan& host, const std::vector<WorkItem>& items, const std::vector<PackDesc>& packs);
cudaError_t fp64_peak(int kind, double* gflops);
cudaError_t ws_profile(unsigned long long out[16], int reset);   // tuning builds (-DFEM2D_WS_PROFILE): cycle counters of k2_ws_kernel

// kernels_fast.cu
cudaError_t launch_k2_sumfact(Plan& plan, uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st, uint32_t* launches);
cudaError_t launch_k2_dmma(Plan& plan, uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st, uint32_t* launches);

// petsc.cu
int device_petsc_prepare(Plan& plan, std::string& err);
int device_petsc_image(Plan& plan, const double* d_vals, void** d_image, uint64_t* bytes, std::string& err);

// kernels_scatter.cu
cudaError_t launch_k3_scatter(const Plan& plan, uint32_t n_ranges, const uint64_t* begins, const uint64_t* ends, double* d_a, double* d_b, int selA, int selB,
                              cudaStream_t st, uint32_t* launches);

}  // namespace fem2d
