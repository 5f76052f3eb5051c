// Host part of the symbolic phase: turns a fem2d_domain_view into blocks / classes / lists / tables / work items
// (plan_types.h).  Integer bookkeeping plus the per-Elem geometry of HierCurlBasisFn::defined_over (basis.rs:365-423),
// evaluated in the reference's operation order -- compile this translation unit with -ffp-contract=off.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fem2d.h"
#include "plan_types.h"

namespace fem2d {

struct HostPlan {
    uint32_t n_elems = 0, n_dofs = 0, i_max = 0, j_max = 0;
    std::vector<uint32_t> bs_off;       // copy of the view's CSR offsets
    std::vector<uint32_t> canon_dof;    // dof id of the k-th function of Elem e in canonical order, at bs_off[e] + k
    std::vector<uint32_t> elem_list;    // list id per Elem
    std::vector<ListDesc> lists;
    std::vector<uint8_t> spec_i, spec_j;
    std::vector<TableDesc> tables;
    std::vector<GramDesc> grams;
    std::vector<ClassDesc> classes;
    std::vector<BlockDesc> blocks;
    std::vector<WorkItem> items;
    std::vector<PackDesc> packs;        // throughput shape: consecutive items the persistent integrator runs together (empty otherwise)
    std::vector<double> elem_dx, elem_dy;   // dx_du, dy_dv per Elem
    uint64_t n_pairs = 0;    // == number of (p,q) integrations the reference performs (x2 for A and B)
    uint64_t n_values = 0;   // entries of V
    uint32_t max_list_n = 0;
    mutable int64_t first_shared_dof = -1;   // smallest DoF carried by more than one Elem (computed on first use by first_shared())
    uint32_t tile_p = 4;            // micro-tile height chosen for the exact integrator (4: throughput, 1: latency)
    uint32_t ws_fold = 0;           // bit 0: every class's uv / vu ratios are powers of two, bit 1: every max(det) is (k2_ws_kernel FOLD)
    uint32_t ws_prod = 1;           // staging warps of the persistent integrator (1 or 2), see K2_WS_TWO_STAGERS_ABOVE
    uint32_t ws_round_slots() const { return (uint32_t)((K2_WS_WARPS - (int)ws_prod) * 32 * K2_WS_TPT); }   // micro-tiles one round of contraction threads holds
    bool use_ws = true;             // throughput shape: big items run in the warp-specialised persistent integrator (FEM2D_K2_WS=0: tuning)
    uint32_t max_slab_stride = 0;   // max over classes of pad4(nU)+pad4(nV) of P (+ the same of Q for non-local classes)
};

// First DoF carried by more than one Elem: the reference numbers all single-Elem (Elem-type) DoFs first (domain.rs:83-96).  Cached in the plan.
uint32_t first_shared(const HostPlan& plan);

// dx_du, dy_dv of every Elem (element.rs:33-50), in the reference's operation order.
int elem_geometry(const fem2d_domain_view* view, std::vector<double>& dx, std::vector<double>& dy, std::string& err);

WorkItem make_item(const HostPlan& plan, uint32_t cls, const std::vector<std::pair<uint32_t, uint32_t>>& ranges, const uint32_t (*cols)[2][2]);

// Integrator launch order of an item list: the items that need K2_THREADS-wide CTAs first (item_is_big: more than K2_SMALL_TILES thread
// slots or a slab row wider than K2_SMALL_STRIDE), each part largest first (longest-processing-time order).
uint32_t item_slab_stride(const HostPlan& plan, const WorkItem& it);
bool item_is_big(const HostPlan& plan, const WorkItem& it);
void order_items(const HostPlan& plan, std::vector<WorkItem>& items);
// Throughput shape with the persistent integrator: groups the items into packs (plan_types.h PackDesc) -- items that fill a round on their
// own stay alone, smaller ones are packed first-fit (largest first) while their micro-tiles fit one round of contraction threads, their
// slab rows stay within K2_PACK_STRIDE functions and their non-local segments share one pair of scaled tables -- and reorders `items` so
// that every pack is contiguous, packs with the most micro-tiles first.  Leaves `packs` empty (and `items` in order_items order) otherwise.
void pack_items(const HostPlan& plan, std::vector<WorkItem>& items, std::vector<PackDesc>& packs);

// Returns FEM2D_OK or a status from include/fem2d.h; err receives a detail message.
int build_host_plan(const fem2d_domain_view* view, bool dedupe, HostPlan& plan, std::string& err);

// Host construction of the pattern (sorted unique keys, first source per slot, extra sources) -- used by device-less
// plans (CPU tests of the host logic / partitioning) and as the cross-check of the device symbolic kernels.
struct HostPattern {
    std::vector<uint32_t> rows, cols, src1;
    std::vector<uint32_t> extra_slot, extra_src;   // 2nd+ contributions, sorted by (slot, generation order)
    uint32_t max_contrib = 0;
};
void build_host_pattern(const HostPlan& plan, HostPattern& pat);

}  // namespace fem2d
