// Plain-data descriptors shared by the host planner (plan_host.cpp) and the CUDA kernels.
//
// Vocabulary (follows the reference's domain):
//   block     one (Elem e, partner d) pair-group of galerkin.rs: d == e  -> "local-local" pairs (galerkin.rs:91-127),
//             d a descendant of e -> "local-desc" pairs (galerkin.rs:138-178).  P = functions of e (sampled over d),
//             Q = functions of d.
//   class     a set of blocks whose inputs are bit-identical (Jacobians, RBS sub-range, materials, BasisSpec sets);
//             integrated once into the value buffer V.
//   list      canonical BasisSpec set of an Elem: sorted by (dir, i, j), U-directed first.
//   table     1-D sampled basis table N, N', T, T' on one axis at GLQ points mapped by x*s + o (basis.rs:372-393).
//   slot      position of a key [row<=col] in the sorted unique upper-triangular pattern.
#pragma once
#include <stdint.h>
#include <string.h>

namespace fem2d {

// Micro-tile of the exact integrator: TP x MT_Q pairs per thread.  TP = 4 is the throughput shape (FP64-issue bound, 16
// independent accumulation chains per thread); TP = 1 is the latency shape used for small / heavily deduplicated plans,
// where 4x more threads with 4x shorter instruction streams fill the machine instead.  The planner picks (HostPlan::tile_p).
// Same-direction pairs (U-U, V-V) accumulate A and B (8 FP64 operations per pair and point); cross-direction pairs (U-V, V-U)
// accumulate A only (3 operations; their B is an exact zero), so a cross-direction tile is twice as wide: TP x MT_QX pairs use
// the same accumulator registers (TP x 4 x {sol, inner}) and cost 48 instead of 64 operations per point at TP = 4 -- warps of
// either kind then take about the same time between two barriers.
constexpr int MT_Q = 2;            // same-direction micro-tile: Q functions (cols) per thread
constexpr int MT_QX = 4;           // cross-direction micro-tile (both shapes; 1 x 2 at TP = 1 measured 16 % slower on cfg 2, equal on cfg 3)
#ifndef FEM2D_TILE_P
#define FEM2D_TILE_P 4             // throughput-shape tile height (tuning builds: -DFEM2D_TILE_P=2)
#endif
#ifndef FEM2D_K2_THREADS
#define FEM2D_K2_THREADS 256
#endif
#ifndef FEM2D_K2_CTAS
#define FEM2D_K2_CTAS 2
#endif
#ifndef FEM2D_K2_ROUNDS
#define FEM2D_K2_ROUNDS 2
#endif
// warp-specialised persistent integrator (kernels_exact.cu k2_ws_kernel): contraction warps + one staging warp per CTA
// K2_WS_WARPS warps per CTA: the first `prod` (1 or 2, chosen per plan: HostPlan::ws_prod) stage slabs, the others contract
#ifndef FEM2D_K2_WS_NBUF
#define FEM2D_K2_WS_NBUF 2
#endif
#ifndef FEM2D_K2_WS_SMEM_KB
#define FEM2D_K2_WS_SMEM_KB 112
#endif
#ifndef FEM2D_K2_WS_TPT
#define FEM2D_K2_WS_TPT 1
#endif
constexpr int K2_WS_TPT = FEM2D_K2_WS_TPT;         // micro-tiles per contraction thread and round (each staged chunk feeds TPT x CONS tiles)
constexpr int K2_WS_WARPS = 8;
constexpr int K2_WS_THREADS = K2_WS_WARPS * 32;
constexpr int K2_WS_MAXREG = (65536 / (FEM2D_K2_CTAS * K2_WS_THREADS)) / 8 * 8 > 128 ? 128 : (65536 / (FEM2D_K2_CTAS * K2_WS_THREADS)) / 8 * 8;   // registers per thread that keep K2_MIN_CTAS CTAs on an SM
constexpr int K2_WS_NBUF = FEM2D_K2_WS_NBUF;       // slab ring depth
constexpr int K2_WS_SMEM_KB = FEM2D_K2_WS_SMEM_KB;
// Staging warps per CTA.  With the roles dealt by SM sub-partition (k2_ws_kernel) two staging warps leave 3 contraction warps + 1 staging
// warp on every sub-partition of an SM, one leaves 4 / 4 / 3 / 3 contraction warps with the fullest sub-partitions pacing every chunk:
// two measured faster on every workload (configs[2] without dedupe 2.14 -> 2.01 ms, cfg4 0.40 -> 0.37 ms, hp1m 8.49 -> 8.23 ms), so the
// threshold (slab columns staged per micro-tile above which a plan gets two) is 0; FEM2D_K2_WS_PROD=1 selects one for tuning.
constexpr double K2_WS_TWO_STAGERS_ABOVE = 0.0;    // staged columns per micro-tile
constexpr int K2_TILE_P = FEM2D_TILE_P;
constexpr int K2_THREADS = FEM2D_K2_THREADS;
constexpr int K2_MIN_CTAS = FEM2D_K2_CTAS;     // CTAs per SM the integrator is compiled for
constexpr int K2_ROUNDS = FEM2D_K2_ROUNDS;    // a work item covers up to K2_ROUNDS * K2_THREADS micro-tiles of one class (slabs staged once)
// Work items with at most K2_SMALL_TILES micro-tiles (small classes: low-order Elems of an hp-mesh, remainders) run in CTAs of
// K2_SMALL_THREADS threads, K2_SMALL_CTAS of which share an SM: a 256-thread CTA with a handful of active threads would hold half
// an SM's registers for the whole quadrature loop.
constexpr int K2_SMALL_THREADS = 64;
constexpr int K2_SMALL_CTAS = 8;
constexpr int K2_SMALL_TILES = 2 * K2_SMALL_THREADS;
constexpr int K2_SMALL_STRIDE = 64;   // ... and a slab row of at most this many functions (1 KB per quadrature point)

struct ListDesc {
    uint32_t off;   // into spec_i / spec_j
    uint32_t n;     // number of functions
    uint32_t nU;    // the first nU are U-directed, the rest V-directed
    uint32_t pad;
};

struct ClassDesc {
    double dxP, dyP, dxQ, dyQ;   // dx_du, dy_dv (element.rs:46-47) of P's Elem and of Q's Elem
    double su, sv;               // P's para_scale (basis.rs:419); (1,1) for local blocks. Q's is always (1,1).
    double eps, mu;              // materials of P's Elem (galerkin.rs:78)
    uint64_t v_off;              // first entry of this class in V; entry (a,b) at v_off + a*nQ + b
    uint32_t listP, listQ;
    uint32_t tabPu, tabPv, tabQu, tabQv;
    uint32_t local;              // 1: d == e (symmetric block, only a <= b is consumed)
    uint32_t n_mt;               // micro-tiles in this class
    uint32_t gramU, gramV;       // ids of the (tabPu,tabQu) / (tabPv,tabQv) 1-D Gram sets (re-ordered modes only)
    ListDesc lp, lq;             // copies of lists[listP], lists[listQ]: one dependent load less in the integrator's prologue
};

// Per-class constants of the exact integrator (Jacobian inverse entries, uv / vu ratios, max(det), 1/mu, eps * glq scales): plan-constant,
// computed once per plan on the device (class_geom_kernel, same expressions as the integrator's prologue) and read by the persistent
// integrator's pack set-up instead of nine FP64 divisions per work item.
struct ClassGeom {
    double jiuP, jivP, jiuQ, jivQ, ratio_uv, ratio_vu, maxdet, coefA, coefB;
};

struct TableDesc {
    double s, o;        // point map x' = x*s + o (glq.rs:238-249)
    uint32_t axis;      // 0: u (orders up to i_max, nu points), 1: v (j_max, nv points)
    uint32_t identity;  // 1: unscaled points (no arithmetic applied, basis.rs:375,392)
};

struct GramDesc {
    uint32_t tabP, tabQ;   // tables of the P side and of the Q side on one axis
    uint32_t axis, pad;
};

// One CTA of the exact integrator: up to ITEM_MAX_RANGES runs of consecutive micro-tiles of one class, plus the slab columns
// (functions) those tiles touch, so that a CTA working on a corner of a class stages only that corner's functions.
constexpr int ITEM_MAX_RANGES = 6;
struct WorkItem {
    uint32_t cls;
    uint32_t mt_count;                    // tiles of all ranges together, <= K2_ROUNDS * K2_THREADS
    uint32_t n_ranges;
    // The first n_same tiles (in range order) are same-direction tiles, the rest cross-direction ones (the class-local numbering
    // puts U-U and V-V first).  Thread slots: the cross-direction tiles start on a warp boundary, so no warp runs both code paths.
    uint32_t n_same;
    uint32_t rbegin[ITEM_MAX_RANGES];     // first micro-tile of each range (class-local numbering)
    uint16_t rcount[ITEM_MAX_RANGES];
    uint16_t stage[2][2][2];              // [side: P, Q][direction group: U, V][begin, end): slab columns to stage
};

// A pack: K2_PACK_MAX or fewer consecutive work items that the persistent integrator processes together (their micro-tiles share the
// contraction threads of one round, their slabs share a ring buffer).  Items with more tiles than one round holds are packs of one.
constexpr int K2_PACK_MAX = 4;
constexpr int K2_PACK_STRIDE = 224;   // widest slab row (functions of all segments, P + Q sides) of a pack of several items
struct PackDesc {
    uint32_t first;   // first work item
    uint32_t n;       // number of work items (segments)
};

struct BlockDesc {
    uint64_t pair_off;   // first pair of this block in the global pair enumeration
    uint32_t cls;
    uint32_t elemP, elemQ;
    uint32_t pad;
};

// ---- micro-tile enumeration (host + device) --------------------------------------------------------------------------------
// Sub-blocks: 0 = U rows x U cols, 1 = U x V, 2 = V x U (non-local only), 3 = V x V.  Class-local tile numbering: the
// same-direction sub-blocks first (0, then 3), then the cross-direction ones (1, then 2).  Local classes use the upper
// triangle (in micro-tile granularity) of the two same-direction sub-blocks.
struct SubBlocks {
    uint32_t cnt[4];
    uint32_t rows[4], cols[4];     // extents
    uint32_t row0[4], col0[4];     // first canonical index
    uint32_t tri[4];
};
#ifdef __CUDACC__
#define FEM2D_HD __host__ __device__
#else
#define FEM2D_HD
#endif
// v = 2^k with |k| <= 64 (v > 0): multiplying by v is exact for every operand the integrator meets, which lets the persistent
// integrator fold such a scale into the quadrature weights (k2_ws_kernel FOLD)
FEM2D_HD inline bool is_pow2_scale(double v) {
    unsigned long long b;
    static_assert(sizeof(b) == sizeof(v), "IEEE-754 binary64");
#ifdef __CUDA_ARCH__
    b = (unsigned long long)__double_as_longlong(v);
#else
    memcpy(&b, &v, sizeof(b));
#endif
    const unsigned e = (unsigned)(b >> 52);   // sign 0 and the biased exponent
    return (b & 0x000fffffffffffffull) == 0ull && e >= 1023u - 64u && e <= 1023u + 64u;
}
FEM2D_HD inline uint32_t mt_div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
FEM2D_HD inline uint32_t mt_sub_at(uint32_t k) { return k == 0 ? 0u : k == 1 ? 3u : k == 2 ? 1u : 2u; }   // k-th sub-block in numbering order
FEM2D_HD inline uint32_t mt_width(uint32_t sub) { return (sub == 1 || sub == 2) ? (uint32_t)MT_QX : (uint32_t)MT_Q; }
FEM2D_HD inline uint32_t mt_tri_count(uint32_t n, uint32_t tp) {
    const uint32_t nrt = mt_div_up(n, tp), nct = mt_div_up(n, MT_Q);
    // tp a multiple of MT_Q: row tile rt starts at column tile rt * (tp / MT_Q) and every row tile keeps at least one column tile, so the
    // sum below is an arithmetic series (the persistent integrator evaluates this per work item on the device)
    if (tp % MT_Q == 0 && nrt > 0) { const uint32_t q = tp / MT_Q; if (nct > (nrt - 1) * q) return nrt * nct - q * (nrt * (nrt - 1) / 2); }
    uint32_t c = 0;
    for (uint32_t rt = 0; rt < nrt; rt++) { const uint32_t lo = rt * tp / MT_Q; if (nct > lo) c += nct - lo; }
    return c;
}
FEM2D_HD inline SubBlocks make_subblocks(uint32_t nP, uint32_t nUP, uint32_t nQ, uint32_t nUQ, uint32_t local, uint32_t tp) {
    SubBlocks s;
    const uint32_t nVP = nP - nUP, nVQ = nQ - nUQ;
    const uint32_t R[4] = {nUP, nUP, nVP, nVP}, Cc[4] = {nUQ, nVQ, nUQ, nVQ};
    const uint32_t r0[4] = {0, 0, nUP, nUP}, c0[4] = {0, nUQ, 0, nUQ};
    for (uint32_t k = 0; k < 4; k++) {
        s.rows[k] = R[k]; s.cols[k] = Cc[k]; s.row0[k] = r0[k]; s.col0[k] = c0[k];
        s.tri[k] = (local && (k == 0 || k == 3)) ? 1u : 0u;
        if (local && k == 2) s.cnt[k] = 0;
        else if (s.tri[k]) s.cnt[k] = mt_tri_count(R[k], tp);
        else s.cnt[k] = mt_div_up(R[k], tp) * mt_div_up(Cc[k], mt_width(k));
    }
    return s;
}
FEM2D_HD inline uint32_t mt_same_count(const SubBlocks& sb) { return sb.cnt[0] + sb.cnt[3]; }   // tiles [0, this) are same-direction

// micro-tile index (class-local) -> sub-block, row tile, column tile
FEM2D_HD inline void decode_tile(const SubBlocks& sb, uint32_t idx, uint32_t tp, uint32_t& sub, uint32_t& rt, uint32_t& ct) {
    uint32_t k = 0;
    sub = mt_sub_at(0);
    while (k < 3 && idx >= sb.cnt[sub]) { idx -= sb.cnt[sub]; k++; sub = mt_sub_at(k); }
    const uint32_t nct = mt_div_up(sb.cols[sub], mt_width(sub));
    if (sb.tri[sub]) {
        rt = 0;
        for (;;) { const uint32_t lo = rt * tp / MT_Q; const uint32_t cnt = nct > lo ? nct - lo : 0; if (idx < cnt) { ct = lo + idx; break; } idx -= cnt; rt++; }
    } else { rt = idx / nct; ct = idx - rt * nct; }
}
// inverse: class-local index of the tile that holds the pair (a, b) (canonical function indices)
FEM2D_HD inline uint32_t encode_tile(const SubBlocks& sb, uint32_t a, uint32_t b, uint32_t nUP, uint32_t nUQ, uint32_t tp) {
    const uint32_t sub = (a < nUP ? 0u : 2u) + (b < nUQ ? 0u : 1u);
    uint32_t idx = 0;
    for (uint32_t k = 0; k < 4 && mt_sub_at(k) != sub; k++) idx += sb.cnt[mt_sub_at(k)];
    const uint32_t w = mt_width(sub);
    const uint32_t rt = (a - sb.row0[sub]) / tp, ct = (b - sb.col0[sub]) / w, nct = mt_div_up(sb.cols[sub], w);
    if (sb.tri[sub]) {
        for (uint32_t q = 0; q < rt; q++) { const uint32_t l0 = q * tp / MT_Q; if (nct > l0) idx += nct - l0; }
        idx += ct - rt * tp / MT_Q;
    } else idx += rt * nct + ct;
    return idx;
}
// thread slots of a work item: its same-direction tiles, padding up to a warp boundary, its cross-direction tiles
FEM2D_HD inline uint32_t item_gap(uint32_t n_same, uint32_t mt_count) { return (n_same == 0 || n_same == mt_count) ? 0u : ((32u - (n_same & 31u)) & 31u); }
FEM2D_HD inline uint32_t item_slots(uint32_t n_same, uint32_t mt_count) { return mt_count + item_gap(n_same, mt_count); }
FEM2D_HD inline uint32_t slab_pad4(uint32_t x) { return (x + 3u) & ~3u; }

}  // namespace fem2d
