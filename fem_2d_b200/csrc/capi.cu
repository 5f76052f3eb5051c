// C-ABI entry points of include/fem2d.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fem2d.h"
#include "device_plan.hpp"

struct fem2d_plan { fem2d::Plan p; };

namespace {
thread_local std::string g_err;
int fail(int status, const std::string& msg) { g_err = msg; return status; }
#define CKS(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? FEM2D_ERR_OUT_OF_MEMORY : FEM2D_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

int device_available(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(FEM2D_ERR_NO_DEVICE, "no CUDA device available: the fem2d numeric path has no CPU fallback"); }
    if (device < 0 || device >= n) return fail(FEM2D_ERR_BAD_ARGUMENT, "device index out of range");
    return FEM2D_OK;
}

int check_numeric_args(const fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode, const double* u_pts, const double* u_w,
                       uint32_t nu, const double* v_pts, const double* v_w, uint32_t nv) {
    if (!plan) return fail(FEM2D_ERR_BAD_ARGUMENT, "null plan");
    // MIN_GLQ_ORDER check (galerkin.rs:51-57) comes first among the numeric arguments
    if (nu < 4 || nv < 4) return fail(FEM2D_ERR_INVALID_GLQ, "Invalid GLQ Settings (the number of GLQ points must be at least 4)");
    if (nu > fem2d::MAX_GLQ || nv > fem2d::MAX_GLQ) return fail(FEM2D_ERR_UNSUPPORTED, "more than 128 GLQ points per axis");
    if (!u_pts || !u_w || !v_pts || !v_w) return fail(FEM2D_ERR_BAD_ARGUMENT, "null GLQ array");
    if (basis_kind != FEM2D_BASIS_HIER_POLY && basis_kind != FEM2D_BASIS_HIER_MAX_ORTHO) return fail(FEM2D_ERR_UNSUPPORTED, "unknown basis space");
    if (basis_kind == FEM2D_BASIS_HIER_MAX_ORTHO && (plan->p.host.i_max > 12 || plan->p.host.j_max > 12))
        return fail(FEM2D_ERR_UNSUPPORTED, "HierMaxOrtho is tabulated up to order 12 (hierarchical_basis_fns.rs:240-252)");
    if ((a_kind != FEM2D_INTEGRAL_CURL_CURL && a_kind != FEM2D_INTEGRAL_L2_INNER) || (b_kind != FEM2D_INTEGRAL_CURL_CURL && b_kind != FEM2D_INTEGRAL_L2_INNER))
        return fail(FEM2D_ERR_UNSUPPORTED, "unknown integral kind");
    if (mode != FEM2D_MODE_EXACT && mode != FEM2D_MODE_SUMFACT && mode != FEM2D_MODE_DMMA) return fail(FEM2D_ERR_UNSUPPORTED, "unknown mode");
    if (plan->p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan: the fem2d numeric path has no CPU fallback");
    return FEM2D_OK;
}
}  // namespace

// Host expansion of a run-compressed index array for the slots [b, e) into out[0 .. e-b), on `threads` host threads.
// start[r] (n_runs + 1 entries, ascending, start[n_runs] = end) is the first slot of run r; slot s of run r gets base(r) + step * (s - start[r]).
//   rows[]:  runs = pattern rows,            start = CSR row offsets,   value = r            (step 0)
//   cols[]:  runs = stretches of consecutive column ids inside a row,   value = first col + offset (step 1)
// The destination is a large (pinned) buffer next to the ones the GPU's DMA engine is filling at the same time, so each worker
// assembles whole 64-byte lines in a small cache-resident block and moves them out with non-temporal stores: no read-for-ownership
// traffic on the memory controllers the DMA writes through.  (Streaming the short runs directly would issue partial-line writes.)
template <class Base>
static void expand_runs(const uint32_t* start, uint64_t n_runs, Base base, uint32_t step, uint64_t b, uint64_t e, uint32_t* out, unsigned threads) {
    if (e <= b) return;
    auto work = [=](uint64_t lo, uint64_t hi) {
        // run holding slot lo: last r with start[r] <= lo
        uint64_t r = (uint64_t)(std::upper_bound(start, start + n_runs + 1, (uint32_t)lo) - start) - 1;
        uint64_t next = start[r + 1];
        constexpr uint32_t BLK = 2048;                          // 8 KB staging block
        alignas(64) uint32_t buf[BLK];
        uint64_t s = lo;
        while (s < hi) {
            uint32_t* dst = out + (s - b);
            // first block: up to the next 64-byte boundary of the destination, then whole blocks
            uint32_t n = (uint32_t)std::min<uint64_t>(BLK, hi - s);
            const uintptr_t mis = reinterpret_cast<uintptr_t>(dst) & 63u;
            if (mis) n = (uint32_t)std::min<uint64_t>(n, (64 - mis) / 4);
            for (uint32_t k = 0; k < n;) {
                while (s + k >= next) { r++; next = start[r + 1]; }
                const uint32_t run = (uint32_t)std::min<uint64_t>(n - k, next - (s + k));
                uint32_t v = base(r) + step * (uint32_t)(s + k - start[r]);
                for (uint32_t q = 0; q < run; q++, v += step) buf[k + q] = v;
                k += run;
            }
#if defined(__SSE2__)
            if (!mis && n % 16 == 0) {
                for (uint32_t k = 0; k < n; k += 4) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k), _mm_load_si128(reinterpret_cast<const __m128i*>(buf + k)));
            } else
#endif
                std::memcpy(dst, buf, (size_t)n * 4);
            s += n;
        }
#if defined(__SSE2__)
        _mm_sfence();   // streaming stores are weakly ordered: drain them before this worker reports completion
#endif
    };
    threads = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(threads, (e - b) / (1u << 20)));
    if (threads == 1) { work(b, e); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++) pool.emplace_back(work, b + (e - b) * t / threads, b + (e - b) * (t + 1) / threads);
    for (auto& th : pool) th.join();
}

extern "C" {

const char* fem2d_version(void) { return "fem2d-b200 0.1.0 (sm_100a)"; }
const char* fem2d_last_error(void) { return g_err.c_str(); }
const char* fem2d_status_string(int s) {
    switch (s) {
        case FEM2D_OK: return "ok";
        case FEM2D_ERR_WRONG_CONTINUITY: return "Wrong Continuity Condition on Domain; Cannot execute Galerkin Sampling!";
        case FEM2D_ERR_EMPTY_DOF_SET: return "No Degrees-of-Freedom Defined over Domain; Cannot execute Galerkin Sampling!";
        case FEM2D_ERR_INVALID_GLQ: return "Invalid GLQ Settings (the number of GLQ points must be at least 4); Cannot execute Galerkin Sampling!";
        case FEM2D_ERR_BAD_ARGUMENT: return "bad argument";
        case FEM2D_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback)";
        case FEM2D_ERR_CUDA: return "CUDA error";
        case FEM2D_ERR_UNSUPPORTED: return "unsupported";
        case FEM2D_ERR_INTERNAL: return "internal error";
        case FEM2D_ERR_OUT_OF_MEMORY: return "out of memory";
    }
    return "unknown status";
}

int fem2d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fem2d_symbolic(const fem2d_domain_view* view, int device, int dedupe, fem2d_plan** out) {
    if (!out) return fail(FEM2D_ERR_BAD_ARGUMENT, "null out");
    *out = nullptr;
    try {
        fem2d_plan* plan = new fem2d_plan();
        std::string err;
        const auto t0 = std::chrono::steady_clock::now();
        int st = fem2d::build_host_plan(view, dedupe != 0, plan->p.host, err);
        const auto t1 = std::chrono::steady_clock::now();
        plan->p.t_host_us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
        if (st != FEM2D_OK) { delete plan; return fail(st, err); }
        if (plan->p.host.n_pairs >= (1ull << 31)) { delete plan; return fail(FEM2D_ERR_UNSUPPORTED, "more than 2^31 pairs"); }
        plan->p.device = device;
        if (device < 0) {
            fem2d::build_host_pattern(plan->p.host, plan->p.host_pattern);
            plan->p.nnz = plan->p.host_pattern.rows.size();
            plan->p.n_extra = plan->p.host_pattern.extra_slot.size();
            plan->p.max_contrib = plan->p.host_pattern.max_contrib;
            uint64_t multi = 0;
            for (size_t k = 0; k < plan->p.host_pattern.extra_slot.size(); k++)
                if (k == 0 || plan->p.host_pattern.extra_slot[k] != plan->p.host_pattern.extra_slot[k - 1]) multi++;
            plan->p.n_multi = multi;
        } else {
            st = device_available(device);
            if (st != FEM2D_OK) { delete plan; return st; }
            st = fem2d::device_symbolic(plan->p, err);
            plan->p.t_device_us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t1).count();
            if (st != FEM2D_OK) { fem2d::device_plan_release(plan->p); delete plan; return fail(st, err); }
        }
        *out = plan;
        return FEM2D_OK;
    } catch (std::bad_alloc&) { return fail(FEM2D_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (std::exception& e) { return fail(FEM2D_ERR_INTERNAL, e.what()); }
}

void fem2d_plan_free(fem2d_plan* plan) {
    if (!plan) return;
    fem2d::device_plan_release(plan->p);
    delete plan;
}

int fem2d_plan_info(const fem2d_plan* plan, uint64_t info[16]) {
    if (!plan || !info) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    std::memset(info, 0, 16 * sizeof(uint64_t));
    const fem2d::Plan& p = plan->p;
    info[0] = p.nnz; info[1] = p.host.n_pairs; info[2] = p.host.blocks.size(); info[3] = p.host.classes.size();
    info[4] = p.host.n_values; info[5] = p.n_multi; info[6] = p.max_contrib; info[7] = p.host.tables.size();
    info[8] = p.host.items.size(); info[9] = p.host.n_dofs; info[10] = p.host.lists.size(); info[11] = p.n_extra;
    info[12] = p.t_host_us; info[13] = p.t_device_us; info[14] = p.host.tile_p; info[15] = p.range_mt_needed;
    return FEM2D_OK;
}

int fem2d_plan_check_work_items(const fem2d_plan* plan, uint64_t out[4]) {
    if (!plan || !out) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    using namespace fem2d;
    const HostPlan& H = plan->p.host;
    const uint32_t tp = H.tile_p;
    uint64_t n_tiles = 0, n_same = 0, n_slots = 0, bad = 0;
    std::vector<std::vector<uint8_t>> seen(H.classes.size());
    for (size_t ci = 0; ci < H.classes.size(); ci++) {
        const ClassDesc& c = H.classes[ci];
        const ListDesc& LP = H.lists[c.listP]; const ListDesc& LQ = H.lists[c.listQ];
        if (std::memcmp(&c.lp, &LP, sizeof(ListDesc)) || std::memcmp(&c.lq, &LQ, sizeof(ListDesc))) bad++;
        const SubBlocks sb = make_subblocks(LP.n, LP.nU, LQ.n, LQ.nU, c.local, tp);
        const uint32_t n_mt = sb.cnt[0] + sb.cnt[1] + sb.cnt[2] + sb.cnt[3];
        if (n_mt != c.n_mt) bad++;
        n_tiles += n_mt; n_same += mt_same_count(sb);
        seen[ci].assign(n_mt, 0);
        // every consumed pair maps to a tile whose decoded extents contain it, and the kind split is where mt_same_count says
        std::vector<uint32_t> cover(n_mt, 0);
        for (uint32_t a = 0; a < LP.n; a++)
            for (uint32_t b = 0; b < LQ.n; b++) {
                const bool u_a = a < LP.nU, u_b = b < LQ.nU;
                if (c.local && ((u_a == u_b && a > b) || (!u_a && u_b))) continue;   // lower triangle of a symmetric block: never read
                const uint32_t idx = encode_tile(sb, a, b, LP.nU, LQ.nU, tp);
                if (idx >= n_mt) { bad++; continue; }
                uint32_t sub, rt, ct;
                decode_tile(sb, idx, tp, sub, rt, ct);
                const uint32_t r0 = sb.row0[sub] + rt * tp, c0 = sb.col0[sub] + ct * mt_width(sub);
                if (a < r0 || a >= r0 + tp || b < c0 || b >= c0 + mt_width(sub)) bad++;
                if ((sub == 0 || sub == 3) != (idx < mt_same_count(sb))) bad++;
                if (sub != (u_a ? 0u : 2u) + (u_b ? 0u : 1u)) bad++;
                cover[idx]++;
            }
        for (uint32_t t = 0; t < n_mt; t++) if (cover[t] == 0) bad++;   // no empty tiles in the numbering
    }
    // launch order: packs cover the item list exactly once, a pack of several items fits one round of contraction threads, stays within
    // K2_PACK_STRIDE functions and holds non-local segments of one table pair only (pack_items); without packs the items that need the
    // wide CTAs form a prefix (order_items)
    if (!H.packs.empty()) {
        uint32_t next = 0;
        for (const PackDesc& pk : H.packs) {
            if (pk.first != next || pk.n == 0 || pk.n > (uint32_t)K2_PACK_MAX || pk.first + pk.n > H.items.size()) { bad++; break; }
            next = pk.first + pk.n;
            if (pk.n > 1) {
                uint32_t same = 0, cross = 0, stride = 0; uint64_t key = ~0ull;
                for (uint32_t k = 0; k < pk.n; k++) {
                    const WorkItem& it = H.items[pk.first + k];
                    if (it.cls >= H.classes.size()) { bad++; continue; }
                    same += it.n_same; cross += it.mt_count - it.n_same; stride += item_slab_stride(H, it);
                    const ClassDesc& c = H.classes[it.cls];
                    if (!c.local) { const uint64_t kk = (uint64_t)c.tabPu << 32 | c.tabPv; if (key != ~0ull && key != kk) bad++; key = kk; }
                }
                if (item_slots(same, same + cross) > H.ws_round_slots() || stride > (uint32_t)K2_PACK_STRIDE) bad++;
            }
        }
        if (next != H.items.size()) bad++;
    }
    bool small_seen = false;
    for (const WorkItem& it : H.items) {
        if (H.packs.empty() && it.cls < H.classes.size()) { const bool big = item_is_big(H, it); if (big && small_seen) bad++; small_seen |= !big; }
        if (it.cls >= H.classes.size() || it.n_ranges == 0 || it.n_ranges > (uint32_t)ITEM_MAX_RANGES) { bad++; continue; }
        const ClassDesc& c = H.classes[it.cls];
        const ListDesc& LP = H.lists[c.listP]; const ListDesc& LQ = H.lists[c.listQ];
        const SubBlocks sb = make_subblocks(LP.n, LP.nU, LQ.n, LQ.nU, c.local, tp);
        const uint32_t same_end = mt_same_count(sb), padUP = slab_pad4(LP.nU), padUQ = slab_pad4(LQ.nU);
        uint32_t cnt = 0, ns = 0, prev_end = 0;
        for (uint32_t r = 0; r < it.n_ranges; r++) {
            if (it.rbegin[r] < prev_end) bad++;                                    // ranges ascend, so same-direction tiles come first
            for (uint32_t t = it.rbegin[r]; t < it.rbegin[r] + it.rcount[r]; t++) {
                if (t >= seen[it.cls].size()) { bad++; continue; }
                seen[it.cls][t]++;
                ns += t < same_end;
                uint32_t sub, rt, ct;
                decode_tile(sb, t, tp, sub, rt, ct);
                // slab columns the tile reads (full tile width: the padding columns must be staged too, as zeros)
                const uint32_t side_q = c.local ? 0u : 1u;
                const uint32_t pr0 = (sub >= 2 ? padUP : 0u) + rt * tp, pr1 = pr0 + tp;
                const uint32_t qc0 = ((sub & 1) ? padUQ : 0u) + ct * mt_width(sub), qc1 = qc0 + mt_width(sub);
                const uint16_t* sp = it.stage[0][sub >= 2]; const uint16_t* sq = it.stage[side_q][sub & 1];
                if (pr0 < sp[0] || std::min(pr1, (sub >= 2 ? padUP + slab_pad4(LP.n - LP.nU) : padUP)) > sp[1]) bad++;
                if (qc0 < sq[0] || std::min(qc1, ((sub & 1) ? padUQ + slab_pad4(LQ.n - LQ.nU) : padUQ)) > sq[1]) bad++;
            }
            cnt += it.rcount[r]; prev_end = it.rbegin[r] + it.rcount[r];
        }
        if (cnt != it.mt_count || ns != it.n_same) bad++;
        n_slots += item_slots(it.n_same, it.mt_count);
    }
    for (auto& s : seen) for (uint8_t k : s) if (k != 1) bad++;
    out[0] = n_tiles; out[1] = n_same; out[2] = n_slots; out[3] = bad;
    return FEM2D_OK;
}

int fem2d_plan_work_info(const fem2d_plan* plan, uint64_t out[8]) {
    if (!plan || !out) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    using namespace fem2d;
    const HostPlan& H = plan->p.host;
    std::memset(out, 0, 8 * sizeof(uint64_t));
    for (const ClassDesc& c : H.classes) {
        const ListDesc& LP = H.lists[c.listP]; const ListDesc& LQ = H.lists[c.listQ];
        const uint64_t uP = LP.nU, vP = LP.n - LP.nU, uQ = LQ.nU, vQ = LQ.n - LQ.nU;
        if (c.local) { out[0] += uP * (uP + 1) / 2 + vP * (vP + 1) / 2; out[1] += uP * vQ; }   // a <= b of a symmetric block (galerkin.rs:91-127)
        else { out[0] += uP * uQ + vP * vQ; out[1] += uP * vQ + vP * uQ; }                      // galerkin.rs:138-178
        const SubBlocks sb = make_subblocks(LP.n, LP.nU, LQ.n, LQ.nU, c.local, H.tile_p);
        out[2] += sb.cnt[0] + sb.cnt[3]; out[3] += sb.cnt[1] + sb.cnt[2];
    }
    out[4] = (uint64_t)H.tile_p * MT_Q; out[5] = (uint64_t)H.tile_p * MT_QX;
    // thread slots the integrator's warps span (whole warps: a pack's same-direction tiles and its cross-direction tiles each round up to 32)
    out[6] = 0;
    if (!H.packs.empty())
        for (const PackDesc& pk : H.packs) {
            uint64_t same = 0, cross = 0;
            for (uint32_t k = 0; k < pk.n; k++) { same += H.items[pk.first + k].n_same; cross += H.items[pk.first + k].mt_count - H.items[pk.first + k].n_same; }
            out[6] += ((same + 31) & ~31ull) + ((cross + 31) & ~31ull);
        }
    else
        for (const WorkItem& it : H.items) out[6] += (((uint64_t)it.n_same + 31) & ~31ull) + (((uint64_t)(it.mt_count - it.n_same) + 31) & ~31ull);
    out[7] = (H.use_ws && H.tile_p == (uint32_t)K2_TILE_P) ? H.ws_prod : 0u;
    return FEM2D_OK;
}

int fem2d_plan_source_map_info(const fem2d_plan* plan, uint64_t info[4]) {
    if (!plan || !info) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    const fem2d::Plan& p = plan->p;
    if (p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan");
    const uint64_t n_chunks = (p.nnz + fem2d::SRC_CHUNK - 1) / fem2d::SRC_CHUNK;
    info[0] = p.n_plain_chunks; info[1] = fem2d::SRC_CHUNK;
    info[2] = n_chunks * 4 + p.nnz * 2 + p.n_plain_chunks * fem2d::SRC_CHUNK * 4;   // bases + 16-bit offsets + the plain chunks' indices
    info[3] = p.nnz * 4;
    return FEM2D_OK;
}

int fem2d_plan_pattern(const fem2d_plan* plan, uint32_t* rows, uint32_t* cols) {
    if (!plan) return fail(FEM2D_ERR_BAD_ARGUMENT, "null plan");
    const fem2d::Plan& p = plan->p;
    if (p.device < 0) {
        if (rows) std::copy(p.host_pattern.rows.begin(), p.host_pattern.rows.end(), rows);
        if (cols) std::copy(p.host_pattern.cols.begin(), p.host_pattern.cols.end(), cols);
        return FEM2D_OK;
    }
    CKS(cudaSetDevice(p.device));
    if (rows) CKS(cudaMemcpy(rows, p.d_rows, p.nnz * 4, cudaMemcpyDeviceToHost));
    if (cols) CKS(cudaMemcpy(cols, p.d_cols, p.nnz * 4, cudaMemcpyDeviceToHost));
    return FEM2D_OK;
}

int fem2d_plan_row_offsets(fem2d_plan* plan, uint64_t* row_ptr) {
    if (!plan || !row_ptr) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    fem2d::Plan& p = plan->p;
    const uint32_t n = p.host.n_dofs;
    if (p.device < 0) {
        const auto& rows = p.host_pattern.rows;
        uint64_t s = 0;
        for (uint32_t r = 0; r <= n; r++) { while (s < rows.size() && rows[s] < r) s++; row_ptr[r] = s; }
        return FEM2D_OK;
    }
    std::string err;
    const int st = fem2d::device_row_ptr_host(p, nullptr, err);
    if (st != FEM2D_OK) return fail(st, err);
    for (uint32_t r = 0; r <= n; r++) row_ptr[r] = p.h_row_ptr[r];
    return FEM2D_OK;
}

int fem2d_plan_pattern_transfer_info(fem2d_plan* plan, uint64_t info[4]) {
    if (!plan || !info) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    fem2d::Plan& p = plan->p;
    if (p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan");
    std::string err;
    const int st = fem2d::device_col_runs_host(p, nullptr, 0, nullptr, nullptr, nullptr, err);
    if (st != FEM2D_OK) return fail(st, err);
    info[0] = ((uint64_t)p.host.n_dofs + 1) * 4; info[1] = p.n_col_runs; info[2] = (2 * p.n_col_runs + 1) * 4; info[3] = p.nnz * 8;
    return FEM2D_OK;
}

int fem2d_plan_pattern_device(const fem2d_plan* plan, const uint32_t** d_rows, const uint32_t** d_cols) {
    if (!plan) return fail(FEM2D_ERR_BAD_ARGUMENT, "null plan");
    if (plan->p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan");
    if (d_rows) *d_rows = plan->p.d_rows;
    if (d_cols) *d_cols = plan->p.d_cols;
    return FEM2D_OK;
}

int fem2d_plan_row_blocks(const fem2d_plan* plan, uint32_t world, uint64_t* bounds) {
    if (!plan || !bounds || world == 0) return fail(FEM2D_ERR_BAD_ARGUMENT, "bad argument");
    const fem2d::Plan& p = plan->p;
    if (p.device < 0) {
        const auto& rows = p.host_pattern.rows;
        bounds[0] = 0; bounds[world] = p.nnz;
        for (uint32_t r = 1; r < world; r++) {
            uint64_t s = p.nnz * r / world;
            while (s < p.nnz && s > 0 && rows[s] == rows[s - 1]) s++;
            bounds[r] = s;
        }
        return FEM2D_OK;
    }
    std::string err;
    int st = fem2d::device_row_block_bounds(p, world, bounds, err);
    return st == FEM2D_OK ? st : fail(st, err);
}

int fem2d_plan_row_blocks_split(const fem2d_plan* plan, uint32_t world, uint64_t* bounds_single, uint64_t* bounds_shared) {
    if (!plan || !bounds_single || !bounds_shared || world == 0) return fail(FEM2D_ERR_BAD_ARGUMENT, "bad argument");
    const fem2d::Plan& p = plan->p;
    // first DoF carried by more than one Elem: the reference numbers all single-Elem (Elem-type) DoFs first (domain.rs:83-96)
    const uint32_t first_shared = fem2d::first_shared(p.host);
    uint64_t split = p.nnz;
    if (p.device < 0) {
        const auto& rows = p.host_pattern.rows;
        split = std::lower_bound(rows.begin(), rows.end(), first_shared) - rows.begin();
        auto part = [&](uint64_t lo, uint64_t hi, uint64_t* b) {
            b[0] = lo; b[world] = hi;
            for (uint32_t r = 1; r < world; r++) {
                uint64_t s = lo + (hi - lo) * r / world;
                while (s < hi && s > lo && rows[s] == rows[s - 1]) s++;
                b[r] = s;
            }
        };
        part(0, split, bounds_single); part(split, p.nnz, bounds_shared);
        return FEM2D_OK;
    }
    // device plan: everything follows from the CSR row offsets (one 4 B x n_dofs copy, cached in the plan; the host-output calls need it anyway)
    std::string err;
    fem2d::Plan& pm = const_cast<fem2d::Plan&>(p);   // caching the pinned copy of the row offsets does not change the plan
    const int st = fem2d::device_row_ptr_host(pm, nullptr, err);
    if (st != FEM2D_OK) return fail(st, err);
    const uint32_t* rp = p.h_row_ptr;
    const uint32_t n = p.host.n_dofs;
    split = first_shared < n ? rp[first_shared] : p.nnz;
    auto part = [&](uint64_t lo, uint64_t hi, uint64_t* b) {
        b[0] = lo; b[world] = hi;
        for (uint32_t r = 1; r < world; r++) {
            uint64_t s = lo + (hi - lo) * r / world;
            if (s > lo && s < hi) {   // advance to the next row start unless s is one
                const uint32_t row = (uint32_t)(std::upper_bound(rp, rp + n + 1, (uint32_t)s) - rp) - 1;   // last row with rp[row] <= s
                if (rp[row] != s) s = std::min<uint64_t>(rp[row + 1], hi);
            }
            b[r] = s;
        }
    };
    part(0, split, bounds_single); part(split, p.nnz, bounds_shared);
    return FEM2D_OK;
}

int fem2d_assemble_device(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode, const double* u_pts, const double* u_w, uint32_t nu,
                          const double* v_pts, const double* v_w, uint32_t nv, uint64_t slot_begin, uint64_t slot_end, double* d_a, double* d_b,
                          void* stream) {
    return fem2d_assemble_device_ranges(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv, 1, &slot_begin, &slot_end, d_a, d_b, stream);
}

int fem2d_assemble_device_ranges(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode, const double* u_pts, const double* u_w, uint32_t nu,
                                 const double* v_pts, const double* v_w, uint32_t nv, uint32_t n_ranges, const uint64_t* slot_begins,
                                 const uint64_t* slot_ends, double* d_a, double* d_b, void* stream) {
    int st = check_numeric_args(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv);
    if (st != FEM2D_OK) return st;
    if (!d_a || !d_b) return fail(FEM2D_ERR_BAD_ARGUMENT, "null output pointer");
    if (n_ranges == 0 || n_ranges > fem2d::MAX_SLOT_RANGES || !slot_begins || !slot_ends) return fail(FEM2D_ERR_BAD_ARGUMENT, "1..4 slot ranges expected");
    fem2d::Plan& p = plan->p;
    cudaStream_t s = (cudaStream_t)stream;
    CKS(cudaSetDevice(p.device));
    // GLQ nodes/weights (inputs of the path, basis.rs:83-90): u_pts | u_w | v_pts | v_w, 128 doubles each
    double h_glq[512];
    std::memset(h_glq, 0, sizeof(h_glq));
    std::copy(u_pts, u_pts + nu, h_glq); std::copy(u_w, u_w + nu, h_glq + 128);
    std::copy(v_pts, v_pts + nv, h_glq + 256); std::copy(v_w, v_w + nv, h_glq + 384);
    // the nodes/weights stay resident on the device: re-upload only when the caller passes different values
    if (!p.glq_valid || std::memcmp(p.h_glq, h_glq, sizeof(h_glq)) != 0) {
        std::memcpy(p.h_glq, h_glq, sizeof(h_glq));
        CKS(cudaMemcpyAsync(p.d_glq, p.h_glq, sizeof(h_glq), cudaMemcpyHostToDevice, s));
        p.glq_valid = true;
    }
    const uint32_t NO = std::max(p.host.i_max, p.host.j_max) + 1, NPT = std::max(nu, nv);
    const size_t tabs_need = p.host.tables.size() * 4 * (size_t)NO * NPT;
    if (tabs_need > p.tabs_capacity) {
        CKS(cudaStreamSynchronize(s));
        fem2d::dev_free(p.d_tabs, s); p.d_tabs = nullptr; p.tabs_capacity = 0;
        CKS(fem2d::dev_malloc((void**)&p.d_tabs, tabs_need * sizeof(double), s));
        p.tabs_capacity = tabs_need;
    }
    if (!p.d_V) CKS(fem2d::dev_malloc((void**)&p.d_V, std::max<uint64_t>(p.host.n_values, 1) * sizeof(double2), s));
    for (int k = 0; k < 4; k++) p.last_launches[k] = 0;
    // Per-phase events are opt-in (fem2d_plan_set_phase_timing): an event between two kernels keeps the second one from starting
    // under programmatic dependent launch, which is how the three kernels of a call overlap their launch latencies.
    const bool timed = p.phase_timing;
    cudaEvent_t* ev = p.ev[p.n_timed_calls % fem2d::Plan::RING];
    if (timed) CKS(cudaEventRecord(ev[0], s));
    CKS(fem2d::launch_k1_tables(p, basis_kind, nu, nv, NO, NPT, s));
    p.last_launches[0] = 1;
    if (timed) CKS(cudaEventRecord(ev[1], s));
    if (mode == FEM2D_MODE_EXACT) {
        const fem2d::WorkItem* items = nullptr; uint32_t n_items = 0;
        const fem2d::PackDesc* packs = nullptr;
        fem2d::ItemSplit split;
        std::string ierr;
        const int ist = fem2d::device_range_items(p, n_ranges, slot_begins, slot_ends, &items, &n_items, &packs, &split, ierr);
        if (ist != FEM2D_OK) return fail(ist, ierr);
        CKS(fem2d::launch_k2_exact(p, items, n_items, packs, split, nu, nv, NO, NPT, s, &p.last_launches[1]));
    }
    else if (mode == FEM2D_MODE_SUMFACT) CKS(fem2d::launch_k2_sumfact(p, nu, nv, NO, NPT, s, &p.last_launches[1]));
    else CKS(fem2d::launch_k2_dmma(p, nu, nv, NO, NPT, s, &p.last_launches[1]));
    if (timed) CKS(cudaEventRecord(ev[2], s));
    CKS(fem2d::launch_k3_scatter(p, n_ranges, slot_begins, slot_ends, d_a, d_b, a_kind == FEM2D_INTEGRAL_L2_INNER, b_kind == FEM2D_INTEGRAL_L2_INNER, s, &p.last_launches[2]));
    if (timed) { CKS(cudaEventRecord(ev[3], s)); p.n_timed_calls++; }
    p.last_launches[3] = p.last_launches[0] + p.last_launches[1] + p.last_launches[2];
    p.n_calls++;
    return FEM2D_OK;
}

int fem2d_plan_set_phase_timing(fem2d_plan* plan, int on) {
    if (!plan) return fail(FEM2D_ERR_BAD_ARGUMENT, "null plan");
    fem2d::Plan& p = plan->p;
    if (p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan");
    if (on && !p.ev[0][0]) {   // the event ring is created on first use
        CKS(cudaSetDevice(p.device));
        for (int r = 0; r < fem2d::Plan::RING; r++) for (int k = 0; k < 4; k++) CKS(cudaEventCreate(&p.ev[r][k]));
    }
    p.phase_timing = on != 0;
    return FEM2D_OK;
}

int fem2d_plan_timing(fem2d_plan* plan, uint32_t calls_back, float ms[4], uint32_t launches[4]) {
    if (!plan) return fail(FEM2D_ERR_BAD_ARGUMENT, "null plan");
    fem2d::Plan& p = plan->p;
    if (p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan");
    if (calls_back >= fem2d::Plan::RING || calls_back >= p.n_timed_calls)
        return fail(FEM2D_ERR_BAD_ARGUMENT, "no timing recorded that far back (fem2d_plan_set_phase_timing enables the per-phase events)");
    cudaEvent_t* ev = p.ev[(p.n_timed_calls - 1 - calls_back) % fem2d::Plan::RING];
    CKS(cudaSetDevice(p.device));
    CKS(cudaEventSynchronize(ev[3]));
    if (ms) {
        for (int k = 0; k < 3; k++) CKS(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
        CKS(cudaEventElapsedTime(&ms[3], ev[0], ev[3]));
    }
    if (launches) std::copy(p.last_launches, p.last_launches + 4, launches);
    return FEM2D_OK;
}
int fem2d_plan_last_timing(fem2d_plan* plan, float ms[4], uint32_t launches[4]) { return fem2d_plan_timing(plan, 0, ms, launches); }

// at_slot: false = the outputs hold the ranges back to back (fem2d_assemble_ranges); true = every slot goes to its own position in outputs
// sized for the whole pattern (several devices of one process fill one set of arrays, fem2d_galerkin_sample_gep_hcurl_multi).
static int assemble_ranges_impl(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode, const double* u_pts, const double* u_w, uint32_t nu,
                                const double* v_pts, const double* v_w, uint32_t nv, uint32_t n_ranges, const uint64_t* slot_begins, const uint64_t* slot_ends,
                                uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals, bool at_slot, unsigned host_threads) {
    int st = check_numeric_args(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv);
    if (st != FEM2D_OK) return st;
    if (!a_vals || !b_vals) return fail(FEM2D_ERR_BAD_ARGUMENT, "null output pointer");
    if (n_ranges == 0 || n_ranges > fem2d::MAX_SLOT_RANGES || !slot_begins || !slot_ends) return fail(FEM2D_ERR_BAD_ARGUMENT, "1..4 slot ranges expected");
    fem2d::Plan& p = plan->p;
    uint64_t b[fem2d::MAX_SLOT_RANGES], e[fem2d::MAX_SLOT_RANGES];
    for (uint32_t k = 0; k < n_ranges; k++) {
        b[k] = slot_begins[k]; e[k] = std::min<uint64_t>(slot_ends[k], p.nnz);
        if (b[k] > e[k]) return fail(FEM2D_ERR_BAD_ARGUMENT, "slot_begin > slot_end");
    }
    CKS(cudaSetDevice(p.device));
    const size_t bytes = std::max<uint64_t>(p.nnz, 1) * sizeof(double);
    if (!p.d_out_a) CKS(fem2d::dev_malloc((void**)&p.d_out_a, bytes));
    if (!p.d_out_b) CKS(fem2d::dev_malloc((void**)&p.d_out_b, bytes));
    // The pattern crosses PCIe in compressed form and is expanded on host threads while the value arrays are in flight: rows[] from
    // the CSR row offsets (4 B per row), cols[] from the runs of consecutive column ids (8 B per run, ~5 runs per row on hp-meshes).
    // They are fetched first: a small copy queued next to the value copies would wait behind them, and so would the expansion.
    uint32_t windows[2 * fem2d::MAX_SLOT_RANGES] = {};
    {
        std::string perr;
        if (rows) { st = fem2d::device_row_ptr_host(p, nullptr, perr); if (st != FEM2D_OK) return fail(st, perr); }
        if (cols) { st = fem2d::device_col_runs_host(p, nullptr, n_ranges, b, e, windows, perr); if (st != FEM2D_OK) return fail(st, perr); }
    }
    st = fem2d_assemble_device_ranges(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv, n_ranges, b, e, p.d_out_a, p.d_out_b, nullptr);
    if (st != FEM2D_OK) return st;
    // D2H of the value slices (outputs hold the ranges back to back); the pattern rides along when requested
    uint64_t off = 0;
    for (uint32_t k = 0; k < n_ranges; k++) {
        const uint64_t n = e[k] - b[k];
        if (n == 0) continue;
        CKS(cudaMemcpyAsync(a_vals + (at_slot ? b[k] : off), p.d_out_a + b[k], n * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
        CKS(cudaMemcpyAsync(b_vals + (at_slot ? b[k] : off), p.d_out_b + b[k], n * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
        off += n;
    }
    if (rows || cols) {   // while the copies above are in flight
        const unsigned hw = std::max(1u, std::min(host_threads, std::thread::hardware_concurrency()));
        const uint32_t* run_col = p.h_col_run_col;
        off = 0;
        for (uint32_t k = 0; k < n_ranges; k++) {
            if (rows) expand_runs(p.h_row_ptr, p.host.n_dofs, [](uint64_t r) { return (uint32_t)r; }, 0u, b[k], e[k], rows + (at_slot ? b[k] : off), hw);
            if (cols && e[k] > b[k]) {   // the window of runs that covers this range (the only part of the run arrays that was fetched)
                const uint32_t lo = windows[2 * k], hi = windows[2 * k + 1];
                const uint32_t* rc = run_col + lo;
                expand_runs(p.h_col_run_slot + lo, (uint64_t)hi - lo, [rc](uint64_t r) { return rc[r]; }, 1u, b[k], e[k], cols + (at_slot ? b[k] : off), hw);
            }
            off += e[k] - b[k];
        }
    }
    CKS(cudaStreamSynchronize(nullptr));
    return FEM2D_OK;
}

int fem2d_assemble_ranges(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode, const double* u_pts, const double* u_w, uint32_t nu,
                          const double* v_pts, const double* v_w, uint32_t nv, uint32_t n_ranges, const uint64_t* slot_begins, const uint64_t* slot_ends,
                          uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals) {
    return assemble_ranges_impl(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv, n_ranges, slot_begins, slot_ends, rows, cols, a_vals, b_vals,
                                false, 8u);
}

int fem2d_assemble_range(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode, const double* u_pts, const double* u_w, uint32_t nu,
                         const double* v_pts, const double* v_w, uint32_t nv, uint64_t slot_begin, uint64_t slot_end,
                         uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals) {
    return fem2d_assemble_ranges(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv, 1, &slot_begin, &slot_end, rows, cols, a_vals, b_vals);
}

int fem2d_assemble(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode, const double* u_pts, const double* u_w, uint32_t nu,
                   const double* v_pts, const double* v_w, uint32_t nv, uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals) {
    return fem2d_assemble_range(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv, 0, UINT64_MAX, rows, cols, a_vals, b_vals);
}

int fem2d_galerkin_sample_gep_hcurl(const fem2d_domain_view* view, int device, int basis_kind, int a_kind, int b_kind, int mode,
                                    const double* u_pts, const double* u_w, uint32_t nu, const double* v_pts, const double* v_w, uint32_t nv,
                                    uint64_t capacity, uint64_t* nnz_out, uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals) {
    if (!view) return fail(FEM2D_ERR_BAD_ARGUMENT, "null view");
    // reference order of the early errors (galerkin.rs:42-59)
    if (view->continuity != FEM2D_CC_HCURL) return fail(FEM2D_ERR_WRONG_CONTINUITY, fem2d_status_string(FEM2D_ERR_WRONG_CONTINUITY));
    if (view->n_dofs == 0) return fail(FEM2D_ERR_EMPTY_DOF_SET, fem2d_status_string(FEM2D_ERR_EMPTY_DOF_SET));
    if (nu < 4 || nv < 4) return fail(FEM2D_ERR_INVALID_GLQ, fem2d_status_string(FEM2D_ERR_INVALID_GLQ));
    fem2d_plan* plan = nullptr;
    int st = fem2d_symbolic(view, device, 1, &plan);
    if (st != FEM2D_OK) return st;
    if (nnz_out) *nnz_out = plan->p.nnz;
    if (plan->p.nnz > capacity) { fem2d_plan_free(plan); return fail(FEM2D_ERR_BAD_ARGUMENT, "output capacity too small (see *nnz_out)"); }
    st = fem2d_assemble(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv, rows, cols, a_vals, b_vals);
    fem2d_plan_free(plan);
    return st;
}

namespace { double g_multi_ms[4 + 4 * 64]; }   // last multi-device call: host plan, then per device symbolic / split / numeric + D2H / release (ms)

// One-shot call on several devices of ONE process: the host half of the symbolic phase runs once, then one host thread per device builds the
// pattern on its device, takes block r of the two-level row partition (fem2d_plan_row_blocks_split), integrates what those rows read and
// copies its slices of A, B (and the expanded rows / cols) to their slot positions in the caller's arrays.  No collective: the <= 2
// contributions of a key are summed on the device that owns its row (halo tiles are recomputed), so the arrays are bit-identical to
// the single-device result whatever the device count.
int fem2d_galerkin_sample_gep_hcurl_multi(const fem2d_domain_view* view, uint32_t n_devices, const int* devices, int basis_kind, int a_kind, int b_kind, int mode,
                                          const double* u_pts, const double* u_w, uint32_t nu, const double* v_pts, const double* v_w, uint32_t nv,
                                          uint64_t capacity, uint64_t* nnz_out, uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals) {
    if (!view) return fail(FEM2D_ERR_BAD_ARGUMENT, "null view");
    // reference order of the early errors (galerkin.rs:42-59)
    if (view->continuity != FEM2D_CC_HCURL) return fail(FEM2D_ERR_WRONG_CONTINUITY, fem2d_status_string(FEM2D_ERR_WRONG_CONTINUITY));
    if (view->n_dofs == 0) return fail(FEM2D_ERR_EMPTY_DOF_SET, fem2d_status_string(FEM2D_ERR_EMPTY_DOF_SET));
    if (nu < 4 || nv < 4) return fail(FEM2D_ERR_INVALID_GLQ, fem2d_status_string(FEM2D_ERR_INVALID_GLQ));
    if (n_devices == 0 || n_devices > 64 || !devices) return fail(FEM2D_ERR_BAD_ARGUMENT, "1..64 devices expected");
    for (uint32_t d = 0; d < n_devices; d++) {
        const int st = device_available(devices[d]);
        if (st != FEM2D_OK) return st;   // (a device index may repeat: its blocks are then assembled one after the other on that device)
    }
    try {
        fem2d::HostPlan host;
        std::string err;
        const auto tm0 = std::chrono::steady_clock::now();
        auto ms_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
        const int hst = fem2d::build_host_plan(view, true, host, err);
        if (hst != FEM2D_OK) return fail(hst, err);
        g_multi_ms[0] = ms_since(tm0);
        if (host.n_pairs >= (1ull << 31)) return fail(FEM2D_ERR_UNSUPPORTED, "more than 2^31 pairs");
        if (n_devices > 1) (void)fem2d::first_shared(host);   // once, before every device plan takes its copy
        std::vector<int> status(n_devices, FEM2D_OK);
        std::vector<std::string> message(n_devices);
        std::vector<uint64_t> nnz(n_devices, 0);
        const unsigned host_threads = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / n_devices));
        auto work = [&](uint32_t r) {
            fem2d_plan* plan = nullptr;
            try {
                plan = new fem2d_plan();
                plan->p.host = host;               // every device plan owns a copy of the (small) host half
                plan->p.device = devices[r];
                std::string e2;
                auto tw = std::chrono::steady_clock::now();
                int st = fem2d::device_symbolic(plan->p, e2);
                g_multi_ms[4 + 4 * r] = ms_since(tw); tw = std::chrono::steady_clock::now();
                if (st == FEM2D_OK) {
                    nnz[r] = plan->p.nnz;
                    if (plan->p.nnz > capacity) { st = FEM2D_ERR_BAD_ARGUMENT; e2 = "output capacity too small (see *nnz_out)"; }
                }
                if (st == FEM2D_OK) {
                    std::vector<uint64_t> b1(n_devices + 1), b2(n_devices + 1);
                    uint64_t begins[2] = {0, 0}, ends[2] = {plan->p.nnz, 0};
                    uint32_t n_ranges = 1;
                    if (n_devices > 1) {
                        st = fem2d_plan_row_blocks_split(plan, n_devices, b1.data(), b2.data());
                        begins[0] = b1[r]; ends[0] = b1[r + 1]; begins[1] = b2[r]; ends[1] = b2[r + 1]; n_ranges = 2;
                    }
                    g_multi_ms[4 + 4 * r + 1] = ms_since(tw); tw = std::chrono::steady_clock::now();
                    if (st == FEM2D_OK)
                        st = assemble_ranges_impl(plan, basis_kind, a_kind, b_kind, mode, u_pts, u_w, nu, v_pts, v_w, nv, n_ranges, begins, ends, rows, cols, a_vals, b_vals,
                                                  true, host_threads);
                    if (st != FEM2D_OK) e2 = g_err;
                    g_multi_ms[4 + 4 * r + 2] = ms_since(tw);
                }
                status[r] = st; message[r] = e2;
            } catch (std::bad_alloc&) { status[r] = FEM2D_ERR_OUT_OF_MEMORY; message[r] = "host allocation failed";
            } catch (std::exception& ex) { status[r] = FEM2D_ERR_INTERNAL; message[r] = ex.what(); }
            const auto tr = std::chrono::steady_clock::now();
            if (plan) { fem2d::device_plan_release(plan->p); delete plan; }
            g_multi_ms[4 + 4 * r + 3] = ms_since(tr);
        };
        if (n_devices == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (uint32_t r = 0; r < n_devices; r++) th.emplace_back(work, r);
            for (auto& t : th) t.join();
        }
        g_multi_ms[1] = ms_since(tm0);
        if (nnz_out) *nnz_out = nnz[0];
        for (uint32_t r = 0; r < n_devices; r++) if (status[r] != FEM2D_OK) return fail(status[r], "device " + std::to_string(devices[r]) + ": " + message[r]);
        return FEM2D_OK;
    } catch (std::bad_alloc&) { return fail(FEM2D_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (std::exception& e) { return fail(FEM2D_ERR_INTERNAL, e.what()); }
}

int fem2d_petsc_aij_size(fem2d_plan* plan, uint64_t* bytes, uint64_t* nnz_full) {
    if (!plan) return fail(FEM2D_ERR_BAD_ARGUMENT, "null plan");
    if (plan->p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan");
    std::string err;
    const int st = fem2d::device_petsc_prepare(plan->p, err);
    if (st != FEM2D_OK) return fail(st, err);
    if (bytes) *bytes = 16 + 4ull * plan->p.host.n_dofs + 12ull * plan->p.aij_nnz_full;
    if (nnz_full) *nnz_full = plan->p.aij_nnz_full;
    return FEM2D_OK;
}

int fem2d_petsc_aij_image(fem2d_plan* plan, const double* d_vals, void* host_image, uint64_t capacity) {
    if (!plan || !d_vals || !host_image) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    if (plan->p.device < 0) return fail(FEM2D_ERR_NO_DEVICE, "host-only plan");
    std::string err;
    void* img = nullptr; uint64_t bytes = 0;
    const int st = fem2d::device_petsc_image(plan->p, d_vals, &img, &bytes, err);
    if (st != FEM2D_OK) return fail(st, err);
    if (bytes > capacity) { fem2d::dev_free(img); return fail(FEM2D_ERR_BAD_ARGUMENT, "image capacity too small (see fem2d_petsc_aij_size)"); }
    const cudaError_t e = cudaMemcpy(host_image, img, bytes, cudaMemcpyDeviceToHost);
    fem2d::dev_free(img);
    if (e != cudaSuccess) return fail(FEM2D_ERR_CUDA, cudaGetErrorString(e));
    return FEM2D_OK;
}

int fem2d_write_petsc_aij(fem2d_plan* plan, const double* d_vals, const char* path) {
    if (!path) return fail(FEM2D_ERR_BAD_ARGUMENT, "null path");
    uint64_t bytes = 0;
    int st = fem2d_petsc_aij_size(plan, &bytes, nullptr);
    if (st != FEM2D_OK) return st;
    void* host = nullptr;
    if (cudaMallocHost(&host, bytes) != cudaSuccess) { cudaGetLastError(); return fail(FEM2D_ERR_OUT_OF_MEMORY, "pinned host allocation failed"); }
    st = fem2d_petsc_aij_image(plan, d_vals, host, bytes);
    if (st == FEM2D_OK) {
        FILE* f = std::fopen(path, "wb");
        if (!f) st = fail(FEM2D_ERR_BAD_ARGUMENT, std::string("cannot open ") + path);
        else {
            if (std::fwrite(host, 1, bytes, f) != bytes) st = fail(FEM2D_ERR_INTERNAL, "short write");
            std::fclose(f);
        }
    }
    cudaFreeHost(host);
    return st;
}

void fem2d_trim_cache(void) { fem2d::dev_cache_trim(); }

void* fem2d_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void fem2d_host_free(void* p) { if (p) cudaFreeHost(p); }

int fem2d_fp64_peak(int device, int kind, double* gflops) {
    if (!gflops) return fail(FEM2D_ERR_BAD_ARGUMENT, "null output");
    int st = device_available(device);
    if (st != FEM2D_OK) return st;
    CKS(cudaSetDevice(device));
    CKS(fem2d::fp64_peak(kind, gflops));
    return FEM2D_OK;
}

/* wall-clock breakdown of the last fem2d_galerkin_sample_gep_hcurl_multi call in ms: out[0] host planner, out[1] whole call, then per device
 * r at out[4 + 4 r ..]: symbolic phase on the device, row-block split, numeric + D2H + host expansion, plan release.  Diagnostic, not in fem2d.h. */
int fem2d_debug_multi_timing(double* out, uint32_t n) {
    for (uint32_t k = 0; k < n && k < 4 + 4 * 64; k++) out[k] = g_multi_ms[k];
    return FEM2D_OK;
}

/* Diagnostic, not in fem2d.h: how full the rounds of the persistent integrator are.  out: micro-tiles, thread slots of all rounds (rounds x
 * contraction threads), packs, rounds, slab columns staged (per round), rounds whose slots are less than half full, HostPlan::ws_fold (scales the
 * integrator folds into the quadrature weights: bit 0 the uv / vu ratios, bit 1 max(det)), 0. */
int fem2d_debug_round_fill(const fem2d_plan* plan, uint64_t out[8]) {
    if (!plan || !out) return fail(FEM2D_ERR_BAD_ARGUMENT, "null argument");
    using namespace fem2d;
    const HostPlan& H = plan->p.host;
    std::memset(out, 0, 8 * sizeof(uint64_t));
    out[6] = H.ws_fold;
    const uint64_t cons = H.ws_round_slots();
    for (const PackDesc& pk : H.packs) {
        uint64_t same = 0, cross = 0, cols = 0;
        for (uint32_t k = 0; k < pk.n; k++) {
            const WorkItem& it = H.items[pk.first + k];
            same += it.n_same; cross += it.mt_count - it.n_same;
            for (int sd = 0; sd < 2; sd++) for (int g = 0; g < 2; g++) cols += it.stage[sd][g][1] - it.stage[sd][g][0];
        }
        const uint64_t slots = same + cross + item_gap((uint32_t)same, (uint32_t)(same + cross));
        const uint64_t rounds = (slots + cons - 1) / cons;
        out[0] += same + cross; out[1] += rounds * cons; out[2] += 1; out[3] += rounds; out[4] += cols * rounds;
        if (2 * (same + cross) < rounds * cons) out[5] += rounds;
        if (std::getenv("FEM2D_DEBUG_FILL")) {   // tiles / slots by kind of pack: 0 one item, several rounds; 1 one item, one round, local; 2 the same, non-local; 3 several items
            static uint64_t cat[4][3];
            const bool nonlocal = !H.classes[H.items[pk.first].cls].local;
            const int k = pk.n > 1 ? 3 : rounds > 1 ? 0 : nonlocal ? 2 : 1;
            cat[k][0] += same + cross; cat[k][1] += rounds * cons; cat[k][2] += cols;
            if (&pk == &H.packs.back())
                for (int q = 0; q < 4; q++) { std::fprintf(stderr, "[fill] kind %d: tiles %llu slots %llu cols %llu\n", q, (unsigned long long)cat[q][0], (unsigned long long)cat[q][1], (unsigned long long)cat[q][2]); cat[q][0] = cat[q][1] = cat[q][2] = 0; }
        }
    }
    return FEM2D_OK;
}

/* tuning builds only (-DFEM2D_WS_PROFILE; zeros otherwise): cycle counters of the warp-specialised integrator; not part of include/fem2d.h */
int fem2d_debug_ws_profile(uint64_t out[16], int reset) {
    unsigned long long t[16];
    CKS(fem2d::ws_profile(t, reset));
    for (int k = 0; k < 16; k++) out[k] = t[k];
    return FEM2D_OK;
}

}  // extern "C"
