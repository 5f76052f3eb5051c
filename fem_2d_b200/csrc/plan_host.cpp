// Host part of the symbolic phase (see plan_host.hpp).  Compile with -ffp-contract=off.
#include "plan_host.hpp"

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <cstring>
#include <numeric>
#include <chrono>
#include <cstdio>
#include <thread>
#include <unordered_map>

namespace fem2d {

namespace {

// 64-bit word-wise mixing hash (keys are memset-padded PODs whose size is a multiple of 4)
inline uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t* p = (const uint8_t*)data;
    size_t k = 0;
    for (; k + 8 <= n; k += 8) {
        uint64_t w; std::memcpy(&w, p + k, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull; h ^= h >> 29;
    }
    for (; k < n; k++) { h ^= p[k]; h *= 1099511628211ull; }
    return h ^ (h >> 32);
}

// Child sub-range (h_refinement.rs:247-279).
inline void sub_range(uint8_t loc, const double in[4], double out[4]) {
    const double mu = (in[0] + in[1]) / 2.0, mv = (in[2] + in[3]) / 2.0;
    const bool west = (loc == FEM2D_LOC_SW || loc == FEM2D_LOC_NW || loc == FEM2D_LOC_W);
    const bool east = (loc == FEM2D_LOC_SE || loc == FEM2D_LOC_NE || loc == FEM2D_LOC_E);
    const bool south = (loc == FEM2D_LOC_SW || loc == FEM2D_LOC_SE || loc == FEM2D_LOC_S);
    const bool north = (loc == FEM2D_LOC_NW || loc == FEM2D_LOC_NE || loc == FEM2D_LOC_N);
    out[0] = east ? mu : in[0];
    out[1] = west ? mu : in[1];
    out[2] = north ? mv : in[2];
    out[3] = south ? mv : in[3];
}
// element.rs:76-78
inline double map_range(double val, double in_min, double in_max, double out_min, double out_max) {
    return (val - in_min) * (out_max - out_min) / (in_max - in_min) + out_min;
}

struct ClassKey {
    double g[10];          // dxP dyP dxQ dyQ su ou sv ov eps mu
    uint32_t listP, listQ, local, pad;
    bool operator==(const ClassKey& o) const { return std::memcmp(this, &o, sizeof(ClassKey)) == 0; }
};
static_assert(sizeof(ClassKey) % 8 == 0, "ClassKey is hashed as 64-bit words");
inline uint64_t class_key_hash(const ClassKey& k) {   // the key is built over a zeroed struct: no padding garbage
    uint64_t w[sizeof(ClassKey) / 8];
    std::memcpy(w, &k, sizeof(ClassKey));
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t x : w) { h = (h ^ x) * 0xff51afd7ed558ccdull; h ^= h >> 32; }
    return h;
}
struct TabKey {
    double s, o; uint32_t axis, identity;
    bool operator==(const TabKey& t) const { return std::memcmp(this, &t, sizeof(TabKey)) == 0; }
};
struct TabKeyHash { size_t operator()(const TabKey& k) const { return (size_t)fnv1a(&k, sizeof(TabKey)); } };

}  // namespace

// Per-Elem constant Jacobian: parametric range of the Elem inside its Element (elem.rs:191-197) mapped to real space
// (element.rs:33-50): diag(dx_du, dy_dv).
int elem_geometry(const fem2d_domain_view* v, std::vector<double>& dx, std::vector<double>& dy, std::string& err) {
    const uint32_t ne = v->n_elems;
    std::vector<double> range(4 * (size_t)ne);
    dx.resize(ne); dy.resize(ne);
    for (uint32_t e = 0; e < ne; e++) {
        const int32_t par = v->elem_parent[e];
        if (par >= (int32_t)e) { err = "elem_parent must precede its children"; return FEM2D_ERR_BAD_ARGUMENT; }
        if (v->elem_element[e] >= v->n_elements) { err = "elem_element out of range"; return FEM2D_ERR_BAD_ARGUMENT; }
        double* r = &range[4 * (size_t)e];
        if (par < 0) { r[0] = -1.0; r[1] = 1.0; r[2] = -1.0; r[3] = 1.0; }
        else {
            if (v->elem_loc[e] > FEM2D_LOC_N) { err = "elem_loc out of range"; return FEM2D_ERR_BAD_ARGUMENT; }
            sub_range(v->elem_loc[e], &range[4 * (size_t)par], r);
        }
        const double* p0 = &v->element_p0[2 * v->elem_element[e]];
        const double* p3 = &v->element_p3[2 * v->elem_element[e]];
        const double real_x_min = map_range(r[0], -1.0, 1.0, p0[0], p3[0]);
        const double real_x_max = map_range(r[1], -1.0, 1.0, p0[0], p3[0]);
        const double real_y_min = map_range(r[2], -1.0, 1.0, p0[1], p3[1]);
        const double real_y_max = map_range(r[3], -1.0, 1.0, p0[1], p3[1]);
        dx[e] = (real_x_max - real_x_min) / 2.0;
        dy[e] = (real_y_max - real_y_min) / 2.0;
    }
    return FEM2D_OK;
}

// Builds a work item from runs of micro-tiles.  `cols` (optional): [side][group][min col, max col + 1) in canonical function
// indices actually touched; nullptr = stage everything.
WorkItem make_item(const HostPlan& P, uint32_t cls, const std::vector<std::pair<uint32_t, uint32_t>>& ranges, const uint32_t (*cols)[2][2]) {
    WorkItem it; std::memset(&it, 0, sizeof(it));
    it.cls = cls; it.n_ranges = (uint32_t)ranges.size();
    const ClassDesc& c = P.classes[cls];
    const ListDesc* L[2] = {&P.lists[c.listP], &P.lists[c.listQ]};
    // ranges are ascending in the class-local numbering, which puts the same-direction tiles first
    const uint32_t same_end = mt_same_count(make_subblocks(L[0]->n, L[0]->nU, L[1]->n, L[1]->nU, c.local, P.tile_p));
    for (size_t k = 0; k < ranges.size(); k++) {
        it.rbegin[k] = ranges[k].first; it.rcount[k] = (uint16_t)ranges[k].second; it.mt_count += ranges[k].second;
        if (ranges[k].first < same_end) it.n_same += std::min(ranges[k].second, same_end - ranges[k].first);
    }
    for (int side = 0; side < 2; side++) {
        const uint32_t nU = L[side]->nU, nV = L[side]->n - nU, padU = slab_pad4(nU);
        uint32_t ub = 0, ue = nU, vb = 0, ve = nV;                      // function ranges inside each direction group
        if (cols) { ub = cols[side][0][0]; ue = cols[side][0][1]; vb = cols[side][1][0]; ve = cols[side][1][1]; }
        // slab columns, rounded outwards to the 4-wide tile rows (the padding columns are staged as zeros)
        it.stage[side][0][0] = (uint16_t)(ue > ub ? (ub & ~3u) : 0); it.stage[side][0][1] = (uint16_t)(ue > ub ? std::min(slab_pad4(ue), padU) : 0);
        it.stage[side][1][0] = (uint16_t)(ve > vb ? padU + (vb & ~3u) : 0); it.stage[side][1][1] = (uint16_t)(ve > vb ? padU + std::min(slab_pad4(ve), slab_pad4(nV)) : 0);
    }
    if (c.local) {   // P and Q share one slab pair: stage the union through the P side
        for (int g = 0; g < 2; g++) {
            const uint16_t pb = it.stage[0][g][0], pe = it.stage[0][g][1], qb = it.stage[1][g][0], qe = it.stage[1][g][1];
            if (qe > qb) { it.stage[0][g][0] = pe > pb ? std::min(pb, qb) : qb; it.stage[0][g][1] = pe > pb ? std::max(pe, qe) : qe; }
            it.stage[1][g][0] = it.stage[1][g][1] = 0;
        }
    }
    return it;
}

uint32_t item_slab_stride(const HostPlan& H, const WorkItem& it) {
    const ClassDesc& c = H.classes[it.cls];
    const ListDesc& LP = H.lists[c.listP]; const ListDesc& LQ = H.lists[c.listQ];
    uint32_t s = slab_pad4(LP.nU) + slab_pad4(LP.n - LP.nU);
    if (!c.local) s += slab_pad4(LQ.nU) + slab_pad4(LQ.n - LQ.nU);
    return s;
}

bool item_is_big(const HostPlan& H, const WorkItem& it) {
    // the latency shape (1 x 2 tiles, small plans) keeps one CTA size
    const uint32_t small_tiles = H.tile_p == 1 ? 0u : (uint32_t)K2_SMALL_TILES;
    return item_slots(it.n_same, it.mt_count) > small_tiles || item_slab_stride(H, it) > (uint32_t)K2_SMALL_STRIDE;
}

void order_items(const HostPlan& H, std::vector<WorkItem>& items) {
    std::stable_sort(items.begin(), items.end(), [](const WorkItem& x, const WorkItem& y) { return x.mt_count > y.mt_count; });
    // the size predicate is not monotone in mt_count (a wide-stride item with few tiles is "big"): partition by the predicate itself
    std::stable_partition(items.begin(), items.end(), [&](const WorkItem& it) { return item_is_big(H, it); });
}

uint32_t first_shared(const HostPlan& H) {
    if (H.first_shared_dof < 0) {
        std::vector<unsigned char> seen(H.n_dofs, 0);
        uint32_t fs = H.n_dofs;
        for (uint32_t d : H.canon_dof) { if (seen[d]) fs = std::min(fs, d); seen[d] = 1; }
        H.first_shared_dof = fs;
    }
    return (uint32_t)H.first_shared_dof;
}

void pack_items(const HostPlan& H, std::vector<WorkItem>& items, std::vector<PackDesc>& packs) {
    packs.clear();
    if (!(H.use_ws && H.tile_p == (uint32_t)K2_TILE_P)) { order_items(H, items); return; }
    const uint32_t round_slots = H.ws_round_slots();
    struct Bin { uint32_t seg[K2_PACK_MAX]; uint32_t n = 0, same = 0, cross = 0, stride = 0; uint64_t key = ~0ull; };
    auto slots_of = [](uint32_t same, uint32_t cross) { return item_slots(same, same + cross); };
    static const bool timing = [] { const char* ev = std::getenv("FEM2D_PLAN_TIMING"); return ev && std::atoi(ev) != 0; }();
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[fem2d planner]   %-26s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    std::vector<uint32_t> order(items.size());
    for (uint32_t k = 0; k < items.size(); k++) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return items[a].mt_count > items[b].mt_count; });
    // slab row and scaled-table key per class, gathered once in class order (the per-item look-ups below then stay in a small array)
    std::vector<uint32_t> cls_stride(H.classes.size());
    std::vector<uint64_t> cls_key(H.classes.size());
    for (size_t c = 0; c < H.classes.size(); c++) {
        const ClassDesc& cd = H.classes[c];
        const ListDesc& LP = H.lists[cd.listP]; const ListDesc& LQ = H.lists[cd.listQ];
        cls_stride[c] = slab_pad4(LP.nU) + slab_pad4(LP.n - LP.nU) + (cd.local ? 0u : slab_pad4(LQ.nU) + slab_pad4(LQ.n - LQ.nU));
        cls_key[c] = cd.local ? ~0ull : ((uint64_t)cd.tabPu << 32 | cd.tabPv);
    }
    uint32_t min_stride = UINT32_MAX;
    for (const WorkItem& it : items) min_stride = std::min(min_stride, cls_stride[it.cls]);
    lap("packs: order");
    // bins hold up to K2_PACK_MAX segments inline; the window of bins that may still take a segment keeps copies of their running sums,
    // so a probe touches 32 contiguous bytes (the first-fit search is most of this function's time on hp-meshes)
    struct Open { uint32_t bin, same, cross, stride, nseg; uint64_t key; };
    std::vector<Bin> bins;
    bins.reserve(items.size());
    std::vector<Open> open;
    for (uint32_t k : order) {
        const WorkItem& it = items[k];
        const uint32_t same = it.n_same, cross = it.mt_count - it.n_same, stride = cls_stride[it.cls];
        const uint64_t key = cls_key[it.cls];
        int slot = -1;
        if (slots_of(same, cross) < round_slots && stride <= (uint32_t)K2_PACK_STRIDE) {
            for (size_t q = 0; q < open.size(); q++) {
                const Open& B = open[q];
                if (B.stride + stride > (uint32_t)K2_PACK_STRIDE) continue;
                if (slots_of(B.same + same, B.cross + cross) > round_slots) continue;
                if (key != ~0ull && B.key != ~0ull && B.key != key) continue;
                slot = (int)q; break;
            }
        }
        uint32_t target;
        if (slot < 0) {
            bins.push_back(Bin());
            target = (uint32_t)bins.size() - 1;
            if (slots_of(same, cross) + 16 < round_slots && stride < (uint32_t)K2_PACK_STRIDE) {
                open.push_back(Open{target, 0, 0, 0, 0, ~0ull});
                if (open.size() > 64) open.erase(open.begin());
                slot = (int)open.size() - 1;
            }
        } else target = open[slot].bin;
        Bin& B = bins[target];
        B.seg[B.n++] = k; B.same += same; B.cross += cross; B.stride += stride;
        if (key != ~0ull) B.key = key;
        if (slot >= 0) {
            Open& o = open[slot];
            o.same = B.same; o.cross = B.cross; o.stride = B.stride; o.nseg = B.n; o.key = B.key;
            // a bin that no later item can join (segment count, slab row, thread slots) leaves the window
            if (o.nseg >= (uint32_t)K2_PACK_MAX || o.stride + min_stride > (uint32_t)K2_PACK_STRIDE || slots_of(o.same, o.cross) + 1 > round_slots)
                open.erase(open.begin() + slot);
        }
    }
    lap("packs: first fit");
    std::vector<uint32_t> bo(bins.size());
    for (uint32_t k = 0; k < bins.size(); k++) bo[k] = k;
    std::stable_sort(bo.begin(), bo.end(), [&](uint32_t a, uint32_t b) { return bins[a].same + bins[a].cross > bins[b].same + bins[b].cross; });
    std::vector<WorkItem> out;
    out.reserve(items.size());
    for (uint32_t b : bo) {
        packs.push_back(PackDesc{(uint32_t)out.size(), bins[b].n});
        for (uint32_t q = 0; q < bins[b].n; q++) out.push_back(items[bins[b].seg[q]]);
    }
    items.swap(out);
    lap("packs: emit");
}

int build_host_plan(const fem2d_domain_view* v, bool dedupe, HostPlan& P, std::string& err) {
    if (!v) { err = "null view"; return FEM2D_ERR_BAD_ARGUMENT; }
    // Reference error order: continuity condition, then empty DoF set (galerkin.rs:42-50).
    if (v->continuity != FEM2D_CC_HCURL) { err = "Wrong Continuity Condition on Domain (required: H(Curl))"; return FEM2D_ERR_WRONG_CONTINUITY; }
    if (v->n_dofs == 0) { err = "No Degrees-of-Freedom Defined over Domain"; return FEM2D_ERR_EMPTY_DOF_SET; }
    const uint32_t ne = v->n_elems;
    if (!v->elem_element || !v->elem_parent || !v->elem_loc || !v->element_p0 || !v->element_p3 || !v->element_eps_re ||
        !v->element_mu_re || !v->bs_off) { err = "null array in view"; return FEM2D_ERR_BAD_ARGUMENT; }
    const uint32_t nbs = v->bs_off[ne];
    if (nbs && (!v->bs_i || !v->bs_j || !v->bs_dir || !v->bs_dof)) { err = "null basis-spec array in view"; return FEM2D_ERR_BAD_ARGUMENT; }
    if (v->i_max > 20 || v->j_max > 20) { err = "expansion order exceeds MAX_POLYNOMIAL_ORDER (20)"; return FEM2D_ERR_UNSUPPORTED; }
    // FEM2D_PLAN_TIMING=1 (tuning): wall time of the planner's phases on stderr
    static const bool timing = [] { const char* ev = std::getenv("FEM2D_PLAN_TIMING"); return ev && std::atoi(ev) != 0; }();
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[fem2d planner] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    P = HostPlan();
    P.n_elems = ne; P.n_dofs = v->n_dofs; P.i_max = v->i_max; P.j_max = v->j_max;
    P.bs_off.assign(v->bs_off, v->bs_off + ne + 1);

    // ---- per-Elem geometry
    if (int st = elem_geometry(v, P.elem_dx, P.elem_dy, err)) return st;

    lap("geometry");
    // ---- canonical BasisSpec lists, pooled by content.
    // Phase A (parallel over Elems): validate, sort each Elem's specs by (dir, i, j), emit canon_dof and a content hash.
    // Phase B (serial): pool identical lists.
    P.canon_dof.resize(nbs);
    P.elem_list.assign(ne, UINT32_MAX);
    for (uint32_t e = 0; e < ne; e++) {
        if (v->bs_off[e + 1] < v->bs_off[e]) { err = "bs_off must be non-decreasing"; return FEM2D_ERR_BAD_ARGUMENT; }
        if (v->bs_off[e + 1] - v->bs_off[e] >= (1u << 11)) { err = "more than 2047 basis specs on one Elem"; return FEM2D_ERR_UNSUPPORTED; }
    }
    std::vector<uint32_t> canon_key(nbs);      // (dir << 16 | i << 8 | j) in canonical order, per Elem at bs_off[e]
    std::vector<uint64_t> elem_hash(ne, 0);
    std::vector<uint32_t> elem_nU(ne, 0);
    // Canonical order of an Elem's functions: U-directed first, then V-directed; inside a direction the Elem-type functions
    // (U: j >= 2, V: i >= 2; basis_spec.rs:49-63) in (i, j) order, then the edge-type functions edge by edge (U: j = 0, j = 1,
    // each by i; V: i = 0, i = 1, each by j).  This follows the reference's DoF numbering (Elem-type DoFs of a leaf are
    // consecutive in generation order, the DoFs of an edge are consecutive: domain.rs:83-149), so consecutive slots of a pattern
    // row read consecutive entries of V (long source runs), and the pairs feeding edge-DoF rows form contiguous micro-tile ranges.
    const uint32_t JSg = v->j_max + 1, KSg = 2 * (v->i_max + 1) * JSg;
    std::vector<uint32_t> rank_cell(KSg), cell_rank(KSg);
    {
        std::vector<std::pair<uint32_t, uint32_t>> keyed;   // (sort key, cell)
        for (uint32_t dir = 0; dir < 2; dir++)
            for (uint32_t i = 0; i <= v->i_max; i++)
                for (uint32_t j = 0; j <= v->j_max; j++) {
                    const uint32_t across = dir == 0 ? j : i;                 // order across the function's tangential edge pair
                    const uint32_t group = across >= 2 ? 0u : 1u + across;   // 0: Elem-type, 1 / 2: the two edges
                    keyed.push_back({dir << 20 | group << 16 | i << 8 | j, (dir * (v->i_max + 1) + i) * JSg + j});
                }
        std::sort(keyed.begin(), keyed.end());
        for (uint32_t r = 0; r < KSg; r++) { rank_cell[r] = keyed[r].second; cell_rank[keyed[r].second] = r; }
    }
    lap("lists: set-up");
    const unsigned n_threads = ne < 4096 ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<int> bad(n_threads, 0);
    auto phase_a = [&](unsigned tid) {
        std::vector<uint32_t> order;
        // Elems are split so that every thread owns about the same number of basis specs
        const uint32_t* off = v->bs_off;
        const uint32_t e0 = (uint32_t)(std::lower_bound(off, off + ne, (uint32_t)((uint64_t)nbs * tid / n_threads)) - off);
        const uint32_t e1 = tid + 1 == n_threads ? ne : (uint32_t)(std::lower_bound(off, off + ne, (uint32_t)((uint64_t)nbs * (tid + 1) / n_threads)) - off);
        // counting placement over the (dir, i, j) key space: O(n + 2 (i_max+1) (j_max+1)) per Elem, no comparison sort
        const uint32_t JS = v->j_max + 1, KS = 2 * (v->i_max + 1) * JS;
        std::vector<int32_t> slot(KS, -1);
        for (uint32_t e = e0; e < e1; e++) {
            const uint32_t b = v->bs_off[e], n = v->bs_off[e + 1] - b;
            if (n == 0) continue;
            order.resize(n);   // packed (dir, i, j, position) keys in canonical order
            bool ok = true, dup = false;
            for (uint32_t k = 0; k < n; k++) {
                const uint32_t dir = v->bs_dir[b + k], i = v->bs_i[b + k], j = v->bs_j[b + k];
                if (dir > 1) { bad[tid] = 1; ok = false; break; }
                if (v->bs_dof[b + k] >= v->n_dofs) { bad[tid] = 2; ok = false; break; }
                if (i > v->i_max || j > v->j_max) { bad[tid] = 3; ok = false; break; }
                int32_t& sl = slot[(dir * (v->i_max + 1) + i) * JS + j];
                if (sl >= 0) dup = true;
                sl = (int32_t)k;
            }
            if (!ok) { std::fill(slot.begin(), slot.end(), -1); continue; }
            if (!dup) {
                uint32_t w = 0;
                for (uint32_t r = 0; r < KS && w < n; r++) {
                    const uint32_t cell = rank_cell[r];
                    if (slot[cell] >= 0) { order[w++] = (uint32_t)slot[cell]; slot[cell] = -1; }
                }
            } else {   // repeated (dir, i, j) on one Elem (never produced by Domain::from_mesh): stable comparison sort keeps all of them
                std::fill(slot.begin(), slot.end(), -1);
                for (uint32_t k = 0; k < n; k++)
                    order[k] = cell_rank[((uint32_t)v->bs_dir[b + k] * (v->i_max + 1) + v->bs_i[b + k]) * JS + v->bs_j[b + k]] << 11 | k;
                std::sort(order.begin(), order.end());
                for (uint32_t k = 0; k < n; k++) order[k] &= 2047u;
            }
            uint32_t nU = 0;
            uint64_t h = 1469598103934665603ull;
            for (uint32_t k = 0; k < n; k++) {
                const uint32_t src = b + order[k];
                const uint32_t key = (uint32_t)v->bs_dir[src] << 16 | (uint32_t)v->bs_i[src] << 8 | v->bs_j[src];
                P.canon_dof[b + k] = v->bs_dof[src];
                canon_key[b + k] = key;
                nU += (key >> 16) == 0;
                h = (h ^ key) * 0x9E3779B97F4A7C15ull; h ^= h >> 29;
            }
            elem_hash[e] = h; elem_nU[e] = nU;
        }
    };
    if (n_threads == 1) phase_a(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < n_threads; t++) th.emplace_back(phase_a, t);
        for (auto& t : th) t.join();
    }
    for (int bcode : bad) {
        if (bcode == 1) { err = "bs_dir must be 0 (U) or 1 (V)"; return FEM2D_ERR_BAD_ARGUMENT; }
        if (bcode == 2) { err = "bs_dof out of range"; return FEM2D_ERR_BAD_ARGUMENT; }
        if (bcode == 3) { err = "basis-spec order exceeds i_max/j_max"; return FEM2D_ERR_BAD_ARGUMENT; }
    }
    lap("lists: canonical order");
    std::unordered_map<uint64_t, std::vector<uint32_t>> list_pool;   // hash -> candidate list ids
    std::vector<uint32_t> list_first_elem;                          // an Elem that carries the list (for content comparison)
    for (uint32_t e = 0; e < ne; e++) {
        const uint32_t b = v->bs_off[e], n = v->bs_off[e + 1] - b;
        if (n == 0) continue;
        const uint32_t nU = elem_nU[e];
        uint32_t id = UINT32_MAX;
        auto& cands = list_pool[elem_hash[e]];
        for (uint32_t cand : cands) {
            const ListDesc& L = P.lists[cand];
            if (L.n != n || L.nU != nU) continue;
            if (std::memcmp(&canon_key[v->bs_off[list_first_elem[cand]]], &canon_key[b], (size_t)n * 4) == 0) { id = cand; break; }
        }
        if (id == UINT32_MAX) {
            id = (uint32_t)P.lists.size();
            P.lists.push_back(ListDesc{(uint32_t)P.spec_i.size(), n, nU, 0});
            list_first_elem.push_back(e);
            for (uint32_t k = 0; k < n; k++) { P.spec_i.push_back((uint8_t)(canon_key[b + k] >> 8)); P.spec_j.push_back((uint8_t)canon_key[b + k]); }
            cands.push_back(id);
        }
        P.elem_list[e] = id;
        P.max_list_n = std::max(P.max_list_n, n);
    }

    lap("lists: pooling");
    // ---- tables: ids 0 / 1 are the unscaled u / v tables
    std::unordered_map<TabKey, uint32_t, TabKeyHash> tab_pool;
    auto table_id = [&](double s, double o, uint32_t axis, uint32_t identity) {
        TabKey k; std::memset(&k, 0, sizeof(k));
        k.s = s; k.o = o; k.axis = axis; k.identity = identity;
        auto it = tab_pool.find(k);
        if (it != tab_pool.end()) return it->second;
        const uint32_t id = (uint32_t)P.tables.size();
        P.tables.push_back(TableDesc{s, o, axis, identity});
        tab_pool.emplace(k, id);
        return id;
    };
    table_id(1.0, 0.0, 0, 1);
    table_id(1.0, 0.0, 1, 1);

    std::unordered_map<uint64_t, uint32_t> gram_pool;
    auto gram_id = [&](uint32_t tp, uint32_t tq, uint32_t axis) {
        const uint64_t k = (uint64_t)tp << 33 | (uint64_t)tq << 1 | axis;
        auto it = gram_pool.find(k);
        if (it != gram_pool.end()) return it->second;
        const uint32_t id = (uint32_t)P.grams.size();
        P.grams.push_back(GramDesc{tp, tq, axis, 0});
        gram_pool.emplace(k, id);
        return id;
    };

    lap("tables");
    // ---- blocks and classes.  Block order: for every Elem d (ascending) its local block, then its blocks with each
    // ancestor that carries functions (nearest ancestor first).
    // Pass 1 (host threads over contiguous Elem ranges): the (ancestor, descendant) pairs that carry functions on both sides and the
    // descendant's sub-range inside the ancestor; pass 2 (serial, in Elem order): class pooling, tables, block numbering.
    struct BlockRec { uint32_t e, d; double su, ou, sv, ov; ClassKey key; uint64_t hash; };   // key without the dedupe-off block number
    auto records_of = [&](uint32_t d0, uint32_t d1, std::vector<BlockRec>& out) {
        std::vector<uint8_t> locs;
        for (uint32_t d = d0; d < d1; d++) {
            if (v->bs_off[d + 1] == v->bs_off[d]) continue;
            locs.clear();
            uint32_t child_on_path = d;
            for (int32_t e = (int32_t)d; e >= 0; e = v->elem_parent[e]) {
                const bool local = (uint32_t)e == d;
                if (!local) locs.push_back(v->elem_loc[child_on_path]);   // locs: from d upwards to the child of e
                child_on_path = (uint32_t)e;
                if (v->bs_off[e + 1] == v->bs_off[e]) continue;
                BlockRec r; std::memset(&r, 0, sizeof(r));
                r.e = (uint32_t)e; r.d = d; r.su = 1.0; r.ou = 0.0; r.sv = 1.0; r.ov = 0.0;
                if (!local) {
                    // relative_parametric_range(e) of d: fold from the child of e down to d (elem.rs:170-188)
                    double rg[4] = {-1.0, 1.0, -1.0, 1.0}, t[4];
                    for (auto it = locs.rbegin(); it != locs.rend(); ++it) { sub_range(*it, rg, t); std::memcpy(rg, t, sizeof(rg)); }
                    r.su = (rg[1] - rg[0]) / 2.0; r.ou = (rg[1] + rg[0]) / 2.0;   // scale_gauss_quad_points glq.rs:238-249
                    r.sv = (rg[3] - rg[2]) / 2.0; r.ov = (rg[3] + rg[2]) / 2.0;
                }
                // class key (the inputs that make two blocks bit-identical) and its hash, computed here in parallel
                const uint32_t elP = v->elem_element[e];
                ClassKey& key = r.key;
                key.g[0] = P.elem_dx[e]; key.g[1] = P.elem_dy[e]; key.g[2] = P.elem_dx[d]; key.g[3] = P.elem_dy[d];
                key.g[4] = r.su; key.g[5] = r.ou; key.g[6] = r.sv; key.g[7] = r.ov;
                key.g[8] = v->element_eps_re[elP]; key.g[9] = v->element_mu_re[elP];
                key.listP = P.elem_list[e]; key.listQ = P.elem_list[d]; key.local = local ? 1u : 0u;
                r.hash = class_key_hash(key);
                out.push_back(r);
            }
        }
    };
    std::vector<std::vector<BlockRec>> rec_parts(n_threads);
    if (n_threads == 1) records_of(0, ne, rec_parts[0]);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < n_threads; t++) th.emplace_back(records_of, (uint32_t)((uint64_t)ne * t / n_threads), (uint32_t)((uint64_t)ne * (t + 1) / n_threads), std::ref(rec_parts[t]));
        for (auto& t : th) t.join();
    }
    lap("blocks: records");
    size_t n_rec = 0;
    for (auto& part : rec_parts) n_rec += part.size();
    P.blocks.reserve(n_rec);
    // class pool: open addressing over the records' precomputed hashes (slot -> class id, keys kept next to the classes); without dedupe
    // every block is its own class and the table is not used
    size_t pool_size = 64;
    while (pool_size < 2 * n_rec) pool_size *= 2;
    std::vector<uint32_t> pool_slot(dedupe ? pool_size : 0, UINT32_MAX);
    std::vector<ClassKey> pool_key;
    if (dedupe) pool_key.reserve(n_rec);
    P.classes.reserve(n_rec);
    for (auto& part : rec_parts)
        for (const BlockRec& r : part) {
            const uint32_t e = r.e, d = r.d;
            const bool local = e == d;
            const uint32_t nd = v->bs_off[d + 1] - v->bs_off[d], nE = v->bs_off[e + 1] - v->bs_off[e];
            const double su = r.su, ou = r.ou, sv = r.sv, ov = r.ov;
            const ClassKey& key = r.key;
            uint32_t cls = UINT32_MAX;
            size_t slot = 0;
            if (dedupe) {
                for (slot = (size_t)r.hash & (pool_size - 1); pool_slot[slot] != UINT32_MAX; slot = (slot + 1) & (pool_size - 1))
                    if (pool_key[pool_slot[slot]] == key) { cls = pool_slot[slot]; break; }
            }
            if (cls == UINT32_MAX) {

                cls = (uint32_t)P.classes.size();
                ClassDesc c; std::memset(&c, 0, sizeof(c));
                c.dxP = key.g[0]; c.dyP = key.g[1]; c.dxQ = key.g[2]; c.dyQ = key.g[3];
                c.su = su; c.sv = sv; c.eps = key.g[8]; c.mu = key.g[9];
                c.listP = key.listP; c.listQ = key.listQ; c.local = key.local;
                c.lp = P.lists[c.listP]; c.lq = P.lists[c.listQ];
                c.tabPu = local ? 0u : table_id(su, ou, 0, 0);
                c.tabPv = local ? 1u : table_id(sv, ov, 1, 0);
                c.tabQu = 0; c.tabQv = 1;
                c.gramU = gram_id(c.tabPu, c.tabQu, 0); c.gramV = gram_id(c.tabPv, c.tabQv, 1);
                c.v_off = P.n_values;
                const ListDesc& LP = P.lists[c.listP]; const ListDesc& LQ = P.lists[c.listQ];
                P.n_values += (uint64_t)LP.n * LQ.n;
                c.n_mt = 0;   // filled once the tile shape is chosen
                P.classes.push_back(c);
                if (dedupe) { pool_slot[slot] = cls; pool_key.push_back(key); }
            }
            BlockDesc b; b.pair_off = P.n_pairs; b.cls = cls; b.elemP = e; b.elemQ = d; b.pad = 0;
            P.blocks.push_back(b);
            P.n_pairs += local ? (uint64_t)nd * (nd + 1) / 2 : (uint64_t)nE * nd;
        }
    if (P.blocks.empty()) { err = "the view carries DoFs but no basis specs: nothing to integrate"; return FEM2D_ERR_BAD_ARGUMENT; }
    if (P.n_values >= (1ull << 31) || P.n_pairs >= (1ull << 32)) { err = "domain too large for 32-bit source indices"; return FEM2D_ERR_UNSUPPORTED; }

    lap("blocks + classes");
    // ---- work items: <= K2_ROUNDS * K2_THREADS micro-tiles each (balanced split); largest classes first (longest-processing-time order)
    std::vector<uint32_t> cls_order(P.classes.size());
    std::iota(cls_order.begin(), cls_order.end(), 0u);
    for (const ClassDesc& c : P.classes) {
        const ListDesc& LP = P.lists[c.listP]; const ListDesc& LQ = P.lists[c.listQ];
        auto p4 = [](uint32_t x) { return (x + 3u) & ~3u; };
        uint32_t s = p4(LP.nU) + p4(LP.n - LP.nU);
        if (!c.local) s += p4(LQ.nU) + p4(LQ.n - LQ.nU);
        P.max_slab_stride = std::max(P.max_slab_stride, s);
    }
    // ---- tile shape: a plan whose 4 x 2 micro-tiles cannot even give every SM one full CTA is latency bound -> 1 x 2 tiles
    auto count_mt = [&](uint32_t tp) {
        uint64_t n = 0;
        for (ClassDesc& c : P.classes) {
            const ListDesc& LP = P.lists[c.listP]; const ListDesc& LQ = P.lists[c.listQ];
            const SubBlocks sb = make_subblocks(LP.n, LP.nU, LQ.n, LQ.nU, c.local, tp);
            c.n_mt = sb.cnt[0] + sb.cnt[1] + sb.cnt[2] + sb.cnt[3];
            n += c.n_mt;
        }
        return n;
    };
    P.tile_p = K2_TILE_P;
    if (count_mt(K2_TILE_P) < (uint64_t)148 * K2_THREADS) { P.tile_p = 1; count_mt(1); }
    lap("items: tile counts");
    {   // largest classes first, ties in class order (a packed key sorts faster than a comparator that chases the 144-byte descriptors)
        std::vector<uint64_t> keyed(P.classes.size());
        for (size_t c = 0; c < P.classes.size(); c++) keyed[c] = (uint64_t)(UINT32_MAX - P.classes[c].n_mt) << 32 | (uint32_t)c;
        std::sort(keyed.begin(), keyed.end());
        for (size_t k = 0; k < keyed.size(); k++) cls_order[k] = (uint32_t)keyed[k];
    }
    // Few, heavily deduplicated classes would leave most of the 148 SMs idle: shrink the item size until there are about two
    // CTAs per SM (each item re-stages its class's slabs, which is cheap next to an idle machine).
    if (const char* ev = std::getenv("FEM2D_K2_WS")) P.use_ws = std::atoi(ev) != 0;   // tuning: 0 = every item in k2_exact_kernel
    if (P.use_ws && P.tile_p == (uint32_t)K2_TILE_P) {
        // staging warps of the persistent integrator: by the slab columns staged per micro-tile (every round of a class stages its columns)
        double cols = 0, tiles = 0;
        for (const ClassDesc& c : P.classes) {
            const ListDesc& LP = P.lists[c.listP]; const ListDesc& LQ = P.lists[c.listQ];
            uint32_t s = slab_pad4(LP.nU) + slab_pad4(LP.n - LP.nU);
            if (!c.local) s += slab_pad4(LQ.nU) + slab_pad4(LQ.n - LQ.nU);
            cols += (double)std::max(1u, (c.n_mt + 223u) / 224u) * s; tiles += c.n_mt;
        }
        P.ws_prod = cols > K2_WS_TWO_STAGERS_ABOVE * tiles ? 2u : 1u;
        if (const char* ev = std::getenv("FEM2D_K2_WS_PROD")) P.ws_prod = std::atoi(ev) == 2 ? 2u : 1u;   // tuning
        // scales that are powers of two in every class (same expressions as class_geom_kernel / the integrator's prologue; this
        // translation unit is compiled with FP contraction off, and IEEE division gives the device's quotients)
        P.ws_fold = 3u;
        for (const ClassDesc& c : P.classes) {
            const double detP = c.dxP * c.dyP - 0.0 * 0.0, detQ = c.dxQ * c.dyQ - 0.0 * 0.0;
            const bool ge = detP >= detQ;
            const double ratio_uv = ge ? c.dxP / c.dyP : c.dxQ / c.dyQ, ratio_vu = ge ? c.dyP / c.dxP : c.dyQ / c.dxQ, maxdet = detP > detQ ? detP : detQ;
            if (!is_pow2_scale(ratio_uv) || !is_pow2_scale(ratio_vu)) P.ws_fold &= ~1u;
            if (!is_pow2_scale(maxdet)) P.ws_fold &= ~2u;
        }
        if (const char* ev = std::getenv("FEM2D_K2_WS_FOLD")) P.ws_fold &= (uint32_t)std::atoi(ev);   // tuning: 0 = never fold
    }
    lap("items: class order, fold");
    uint32_t cap = K2_ROUNDS * (P.use_ws && P.tile_p == (uint32_t)K2_TILE_P ? P.ws_round_slots() / K2_WS_TPT : (uint32_t)K2_THREADS);
    const uint32_t min_cap = P.tile_p == 1 ? 256u : 64u;   // latency shape: one full round per CTA measured best (128: +10 %, 512: +30 %)
    for (; cap > min_cap; cap /= 2) {
        uint64_t n = 0;
        for (const ClassDesc& c : P.classes) n += (c.n_mt + cap - 1) / cap;
        if (n >= 2 * 148) break;
    }
    // (classes are independent: on plans with many classes the list is cut into contiguous pieces of cls_order, one per host thread,
    // and the pieces are concatenated in order, so the item list does not depend on the thread count)
    auto items_of = [&](size_t k0, size_t k1, std::vector<WorkItem>& out) {
        for (size_t kc = k0; kc < k1; kc++) {
            const uint32_t c = cls_order[kc];
            const uint32_t n_mt = P.classes[c].n_mt;
            const ClassDesc& cd = P.classes[c];
            const ListDesc& LP = P.lists[cd.listP]; const ListDesc& LQ = P.lists[cd.listQ];
            const uint32_t same_end = mt_same_count(make_subblocks(LP.n, LP.nU, LQ.n, LQ.nU, cd.local, P.tile_p));
            // equal shares; one more item if the warp alignment of the cross-direction tiles would push a share past the cap
            uint32_t n_items = (n_mt + cap - 1) / cap;
            for (;; n_items++) {
                bool fits = true;
                for (uint32_t k = 0; k < n_items && fits; k++) {
                    const uint32_t b = (uint32_t)((uint64_t)n_mt * k / n_items), e = (uint32_t)((uint64_t)n_mt * (k + 1) / n_items);
                    const uint32_t ns = b < same_end ? std::min(e, same_end) - b : 0u;
                    fits = item_slots(ns, e - b) <= cap;
                }
                if (fits || n_items >= n_mt) break;
            }
            for (uint32_t k = 0; k < n_items; k++) {
                const uint32_t b = (uint32_t)((uint64_t)n_mt * k / n_items), e = (uint32_t)((uint64_t)n_mt * (k + 1) / n_items);
                if (e > b) out.push_back(make_item(P, c, {{b, e - b}}, nullptr));
            }
        }
    };
    lap("items: cap");
    const unsigned item_threads = cls_order.size() < 8192 ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (item_threads == 1) items_of(0, cls_order.size(), P.items);
    else {
        // pieces of about equal item counts (cls_order has the largest classes first: equal class counts would give the first thread most of the items)
        lap("items: pre");
        std::vector<uint64_t> before(cls_order.size() + 1, 0);
        for (size_t k = 0; k < cls_order.size(); k++) before[k + 1] = before[k] + (P.classes[cls_order[k]].n_mt + cap - 1) / cap;
        std::vector<size_t> cut(item_threads + 1, cls_order.size());
        cut[0] = 0;
        for (unsigned t = 1; t < item_threads; t++)
            cut[t] = (size_t)(std::lower_bound(before.begin(), before.end(), before.back() * t / item_threads) - before.begin());
        std::vector<std::vector<WorkItem>> parts(item_threads);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < item_threads; t++) {
            parts[t].reserve((size_t)(before[cut[t + 1]] - before[cut[t]]) + 16);
            th.emplace_back(items_of, cut[t], cut[t + 1], std::ref(parts[t]));
        }
        for (auto& t : th) t.join();
        lap("items: threads");
        P.items.reserve((size_t)before.back() + 16 * item_threads);
        for (auto& part : parts) P.items.insert(P.items.end(), part.begin(), part.end());
    }
    lap("work items");
    pack_items(P, P.items, P.packs);
    lap("packs");
    return FEM2D_OK;
}

void build_host_pattern(const HostPlan& P, HostPattern& pat) {
    struct Rec { uint64_t key; uint32_t src; };
    std::vector<Rec> recs;
    recs.reserve(P.n_pairs);
    for (const BlockDesc& b : P.blocks) {
        const ClassDesc& c = P.classes[b.cls];
        const uint32_t nP = P.lists[c.listP].n, nQ = P.lists[c.listQ].n;
        const uint32_t* dp = &P.canon_dof[P.bs_off[b.elemP]];
        const uint32_t* dq = &P.canon_dof[P.bs_off[b.elemQ]];
        for (uint32_t a = 0; a < nP; a++)
            for (uint32_t q = c.local ? a : 0; q < nQ; q++) {
                const uint32_t r = std::min(dp[a], dq[q]), cc = std::max(dp[a], dq[q]);
                recs.push_back(Rec{(uint64_t)r << 32 | cc, (uint32_t)(c.v_off + (uint64_t)a * nQ + q)});
            }
    }
    std::stable_sort(recs.begin(), recs.end(), [](const Rec& x, const Rec& y) { return x.key < y.key; });
    pat = HostPattern();
    uint32_t run = 0;
    for (size_t k = 0; k < recs.size(); k++) {
        if (k == 0 || recs[k].key != recs[k - 1].key) {
            pat.rows.push_back((uint32_t)(recs[k].key >> 32)); pat.cols.push_back((uint32_t)recs[k].key); pat.src1.push_back(recs[k].src);
            run = 1;
        } else {
            pat.extra_slot.push_back((uint32_t)pat.rows.size() - 1); pat.extra_src.push_back(recs[k].src);
            run++;
        }
        pat.max_contrib = std::max(pat.max_contrib, run);
    }
}

}  // namespace fem2d
