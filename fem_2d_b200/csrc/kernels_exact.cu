// K1 (basis-table sampler) and K2-exact (per-pair integrator that replays the reference's floating-point order).
//
// THIS FILE MUST BE COMPILED WITH -fmad=false: every product/sum below is a separately rounded IEEE-754 operation, in
// the evaluation order of the reference expressions cited next to it.  Results are then bit-identical to the reference
// algorithm (same GLQ nodes in).  Citations are relative to /root/reference/.
//
// Data flow of one pair (p on P, q on Q), per quadrature point (m, n)   [SURVEY.md App. A.4]:
//   U-directed f:  curl_f = -(((jinv.u[0] * (N_i(m) * T'_j(n))) * ps_other[0]))        basis.rs:235-242, integrals.rs:43-48,240
//                  val_f  =  (jinv.u[0] * N_i(m)) * T_j(n)                              basis.rs:225-227
//   V-directed f:  curl_f =  ((jinv.v[1] * (T'_i(m) * N_j(n))) * ps_other[1])          basis.rs:245-252
//                  val_f  =  (jinv.v[1] * T_i(m)) * N_j(n)                              basis.rs:230-232
//   (the other vector component is (-0.0 * x): it only ever adds a signed zero, which cannot change a non-zero sum nor the
//    +0.0 the accumulators start from, so dropping it is bit-exact)
//   A: inner += ((curl_p * curl_q) [* ratio]) * v_w[n];  sol += inner * u_w[m];  A = (1/mu) * sol      integrals.rs:36-91, glq.rs:19-32
//   B: inner += ((val_p * val_q) * max(det_P, det_Q)) * v_w[n]; ...; B = ((eps * glq_P) * glq_Q) * sol  integrals.rs:302-353
// curl_f and val_f depend on one function only -> they are staged per block in shared memory ("slabs"), and each thread
// contracts a register tile of pairs over the points in strict (m outer, n inner) order: 4 x 2 same-direction pairs (A and B)
// or 4 x 4 cross-direction pairs (A only; B is an exact zero) -- the same accumulator registers and about the same work.
#include <cuda_runtime.h>

#include <mutex>

#include "basis_device.cuh"
#include "device_plan.hpp"

namespace fem2d {

namespace {

// ---------------------------------------------------------------------------------------------------------------- K1
// One CTA per table, one thread per point.  Output layout: out[((arr * NO) + order) * NPT + point], arr: 0 N, 1 N', 2 T, 3 T'.
__global__ void k1_tables_kernel(const TableDesc* __restrict__ tabs, double* __restrict__ out, uint32_t NO, uint32_t NPT,
                                 const double* __restrict__ glq, uint32_t nu, uint32_t nv, uint32_t i_max, uint32_t j_max, int basis,
                                 uint32_t* __restrict__ work_counter) {
    // the persistent integrator (k2_ws_kernel) hands out its work items through this counter; every numeric call starts at 0
    if (blockIdx.x == 0 && threadIdx.x == 0) *work_counter = 0u;
    // programmatic dependent launch: let the integrator's CTAs start their prologue (work item, class, tile decode) right away;
    // they wait for this grid's completion (cudaGridDependencySynchronize) before they read the tables
    cudaTriggerProgrammaticLaunchCompletion();
    const TableDesc t = tabs[blockIdx.x];
    const uint32_t np = t.axis ? nv : nu, nmax = t.axis ? j_max : i_max;
    const double* pts = glq + (t.axis ? 256 : 0);
    double* o = out + (size_t)blockIdx.x * 4 * NO * NPT;
    for (uint32_t p = threadIdx.x; p < np; p += blockDim.x) {
        // RBS ancestor -> descendant point map (basis.rs:372-393, glq.rs:238-249)
        const double x = t.identity ? pts[p] : pts[p] * t.s + t.o;
        basis_at_point(basis, nmax, x, [&](int arr, uint32_t n, double val) { o[((size_t)arr * NO + n) * NPT + p] = val; });
    }
}

// ---------------------------------------------------------------------------------------------------------------- K2
struct K2Args {
    const ClassDesc* classes; const ListDesc* lists; const uint8_t* spec_i; const uint8_t* spec_j;
    const WorkItem* items; const double* tabs; const double* glq; double2* V;
    uint32_t NO, NPT, nu, nv, slab_doubles;   // slab_doubles: shared-memory doubles available for the slabs of one CTA
    // 1: this grid follows the sampler in the stream (it waits for it before reading the tables and only then lets its dependents
    // launch); 0: it follows the first integrator grid, which it does not depend on -- it runs alongside it and waits for it only
    // before exiting, so that a grid waiting on this one has transitively waited on both.
    int follows_sampler;
    uint32_t col_cap;    // warp-specialised integrator: entries of its column table (>= the widest pack's slab row, multiple of 4)
    const ClassGeom* geom;   // warp-specialised integrator: per-class constants (class_geom_kernel)
    int roles_by_subpartition;   // warp-specialised integrator: 1 = staging warps chosen by SM sub-partition (default), 0 = by warp index (tuning)
};

__device__ __forceinline__ uint32_t pad4(uint32_t x) { return (x + 3u) & ~3u; }

// Row stride (doubles) of a sampled table inside the staging warps' shared-memory cache.  The lanes of a staging warp read the rows of
// 32 different (i, j) orders at once; with the natural stride (8 or 12 points = 64 or 96 bytes) those rows start in 2 or 4 bank
// groups only.  An odd stride puts up to 16 different orders into 16 different 8-byte bank pairs.
#ifndef FEM2D_K2_WS_TABPAD
#define FEM2D_K2_WS_TABPAD 1
#endif
__host__ __device__ __forceinline__ uint32_t ws_tab_stride(uint32_t npt) { return FEM2D_K2_WS_TABPAD ? (npt | 1u) : npt; }

// Accumulators of one thread: acc[2][TP][MT_Q].  Same-direction tile (TP x MT_Q pairs): [0] = A, [1] = B.  Cross-direction tile
// (TP x MT_QX pairs, A only): column c lives in [c >> 1][r][c & 1].
template <int TP>
__device__ __forceinline__ void load_rows(const double* __restrict__ p, double (&v)[TP]) {
    if (TP == 1) v[0] = p[0];
    else {
#pragma unroll
        for (int r = 0; r + 1 < TP; r += 2) { const double2 t = *reinterpret_cast<const double2*>(p + r); v[r] = t.x; v[r + 1] = t.y; }
    }
}

// One point of a same-direction tile.  The operations of one pair form a dependent chain (8-cycle FP64 latency each); they are
// written stage by stage over the independent pairs so that the in-order issue always has independent work (same operations,
// same order per value).
template <int TP>
__device__ __forceinline__ void contract_point_same(const double* __restrict__ cp, const double* __restrict__ cq, const double* __restrict__ fp,
                                                    const double* __restrict__ fq, double ratio, double maxdet, double w,
                                                    double (&in)[2][TP][MT_Q]) {
    double pc[TP], qc[MT_Q], pf[TP], qf[MT_Q], tA[TP][MT_Q], tB[TP][MT_Q];
    load_rows<TP>(cp, pc); load_rows<MT_Q>(cq, qc);
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) tA[r][c] = pc[r] * qc[c];              // p_curl * q_curl
    load_rows<TP>(fp, pf); load_rows<MT_Q>(fq, qf);
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) tB[r][c] = pf[r] * qf[c];              // V2D::dot(f_p, f_q): the second product is a signed zero
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) tA[r][c] = tA[r][c] * ratio;           // * max_uv_ratios / max_vu_ratios (integrals.rs:50,86)
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) tB[r][c] = tB[r][c] * maxdet;          // * partial_max(det_P, det_Q) (integrals.rs:312-315)
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) tA[r][c] = tA[r][c] * w;
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) tB[r][c] = tB[r][c] * w;
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) in[0][r][c] = in[0][r][c] + tA[r][c];  // inner_solution += integrand * v_w (glq.rs:27)
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) in[1][r][c] = in[1][r][c] + tB[r][c];
}

// One point of a cross-direction tile: (p_curl * q_curl) * v_w only (integrals.rs:53-84: no ratio factor; the mass integrand is
// a signed zero).
template <int TP>
__device__ __forceinline__ void contract_point_cross(const double* __restrict__ cp, const double* __restrict__ cq, double w,
                                                     double (&in)[2][TP][MT_Q]) {
    constexpr int XW = MT_QX;
    double pc[TP], qc[XW], tA[TP][XW];
    load_rows<TP>(cp, pc); load_rows<XW>(cq, qc);
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < XW; c++) tA[r][c] = pc[r] * qc[c];
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < XW; c++) tA[r][c] = tA[r][c] * w;
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < XW; c++) in[c >> 1][r][c & 1] = in[c >> 1][r][c & 1] + tA[r][c];
}

// Contract `run` consecutive points of one quadrature row (n .. n+run-1 of row m) into the inner accumulators.
template <bool SAME, int TP>
__device__ __forceinline__ void contract_run(const double*& cp, const double*& cq, const double*& fp, const double*& fq, uint32_t strideP,
                                             uint32_t strideQ, const double* __restrict__ vw, uint32_t run, double ratio, double maxdet,
                                             double (&in)[2][TP][MT_Q]) {
#pragma unroll 4
    for (uint32_t k = 0; k < run; k++) {
        if (SAME) contract_point_same<TP>(cp, cq, fp, fq, ratio, maxdet, vw[k], in);
        else contract_point_cross<TP>(cp, cq, vw[k], in);
        cp += strideP; cq += strideQ;
        if (SAME) { fp += strideP; fq += strideQ; }
    }
}

template <int TP, int NT>
__global__ void __launch_bounds__(NT, NT == K2_THREADS ? K2_MIN_CTAS : K2_SMALL_CTAS) k2_exact_kernel(const K2Args g) {
    extern __shared__ __align__(16) double smem[];
    // programmatic dependent launch: the next grid (second integrator grid or the scatter kernel, which starts by loading its
    // source offsets and waits before reading V) may start once every CTA of this one has passed this point
    if (!g.follows_sampler) cudaTriggerProgrammaticLaunchCompletion();
    const WorkItem it = g.items[blockIdx.x];
    const ClassDesc c = g.classes[it.cls];
    const ListDesc LP = c.lp, LQ = c.lq;
    const uint32_t nP = LP.n, nUP = LP.nU, nQ = LQ.n, nUQ = LQ.nU;
    const uint32_t strideP = pad4(nUP) + pad4(nP - nUP);
    const uint32_t strideQ = c.local ? strideP : pad4(nUQ) + pad4(nQ - nUQ);
    const uint32_t nu = g.nu, nv = g.nv, npts = nu * nv;
    // points per staging chunk: as many as this class's slab rows (C and F of P, and of Q unless local) fit
    const uint32_t chunk = min(npts, g.slab_doubles / (2 * (strideP + (c.local ? 0u : strideQ))));

    double* s_uw = smem;                       // [128]
    double* s_vw = smem + 128;                 // [128]
    double* s_CP = smem + 256;                 // [chunk][strideP]
    double* s_FP = s_CP + (size_t)chunk * strideP;
    double* s_CQ = c.local ? s_CP : s_FP + (size_t)chunk * strideP;
    double* s_FQ = c.local ? s_FP : s_CQ + (size_t)chunk * strideQ;
    for (uint32_t k = threadIdx.x; k < nu; k += blockDim.x) s_uw[k] = g.glq[128 + k];
    for (uint32_t k = threadIdx.x; k < nv; k += blockDim.x) s_vw[k] = g.glq[384 + k];

    // ---- per-class constants (HierCurlBasisFn::defined_over, basis.rs:395-413; M2D::det / inverse, space.rs:138-147).
    // An FP64 division is a ~25-instruction dependent chain: the nine quotients are computed once per CTA, one per lane of the
    // second warp (the first one builds the tile enumeration meanwhile), and read back after the barrier.
    const double detP = c.dxP * c.dyP - 0.0 * 0.0, detQ = c.dxQ * c.dyQ - 0.0 * 0.0;
    __shared__ double s_quot[9];
    __shared__ SubBlocks sb;
    if (threadIdx.x == 0) sb = make_subblocks(nP, nUP, nQ, nUQ, c.local, TP);
    if (threadIdx.x >= 32 && threadIdx.x < 41) {
        const uint32_t k = threadIdx.x - 32;
        const double num = k == 0 ? c.dyP : k == 1 ? c.dxP : k == 2 ? c.dyQ : k == 3 ? c.dxQ : k == 4 ? c.dxP : k == 5 ? c.dxQ : k == 6 ? c.dyP : k == 7 ? c.dyQ : 1.0;
        const double den = k < 2 ? detP : k < 4 ? detQ : k == 4 ? c.dyP : k == 5 ? c.dyQ : k == 6 ? c.dxP : k == 7 ? c.dxQ : c.mu;
        s_quot[k] = num / den;
    }
    __syncthreads();
    const double jiuP = s_quot[0], jivP = s_quot[1];           // jac_inv.u[0] = dy_dv / det, jac_inv.v[1] = dx_du / det
    const double jiuQ = s_quot[2], jivQ = s_quot[3];
    const double ge = (double)(detP >= detQ), lt = (double)(detP < detQ);
    const double ratio_uv = ge * s_quot[4] + lt * s_quot[5];   // max_uv_ratios integrals.rs:250-259, basis.rs:341-343: ge * (dxP / dyP) + lt * (dxQ / dyQ)
    const double ratio_vu = ge * s_quot[6] + lt * s_quot[7];   // max_vu_ratios integrals.rs:262-271, basis.rs:346-348: ge * (dyP / dxP) + lt * (dyQ / dxQ)
    const double maxdet = detP > detQ ? detP : detQ;           // partial_max integrals.rs:421-423
    const double coefA = s_quot[8];                            // 1.0 / mu, integrals.rs:37
    const double coefB = c.eps * (c.su * c.sv) * (1.0 * 1.0);  // eps * p.glq_scale() * q.glq_scale() integrals.rs:303-305

    const uint32_t AS = g.NO * g.NPT;   // stride between the four arrays N, N', T, T'
    const double* tPu = g.tabs + (size_t)c.tabPu * 4 * AS;
    const double* tPv = g.tabs + (size_t)c.tabPv * 4 * AS;
    const double* tQu = g.tabs + (size_t)c.tabQu * 4 * AS;
    const double* tQv = g.tabs + (size_t)c.tabQv * 4 * AS;
    const bool single_chunk = chunk >= npts;
    double2* out = g.V + c.v_off;
    if (g.follows_sampler) {
        cudaGridDependencySynchronize();   // the sampler's tables (and, transitively, the previous call's readers of V) are complete
        cudaTriggerProgrammaticLaunchCompletion();
    }

    // thread slots: same-direction tiles, padding to a warp boundary, cross-direction tiles (plan_types.h item_slots)
    const uint32_t gap = item_gap(it.n_same, it.mt_count), n_slots = it.mt_count + gap;
    for (uint32_t round0 = 0; round0 < n_slots; round0 += NT) {
        // ---- my micro-tile of this round
        const uint32_t slot = round0 + threadIdx.x;
        const bool active = slot < n_slots && !(slot >= it.n_same && slot < it.n_same + gap);
        uint32_t sub = 0, row0 = 0, col0 = 0, row_end = 0, col_end = 0, prow = 0, pcol = 0;
        if (active) {
            uint32_t li = slot < it.n_same ? slot : slot - gap, r = 0;
            while (li >= it.rcount[r]) { li -= it.rcount[r]; r++; }        // which of the item's tile ranges
            uint32_t rt, ct;
            decode_tile(sb, it.rbegin[r] + li, TP, sub, rt, ct);
            row0 = sb.row0[sub] + rt * TP; col0 = sb.col0[sub] + ct * mt_width(sub);
            row_end = sb.row0[sub] + sb.rows[sub]; col_end = sb.col0[sub] + sb.cols[sub];
            prow = (sub >= 2 ? pad4(nUP) - nUP : 0) + row0;            // slab column of the canonical row index
            pcol = ((sub & 1) ? pad4(nUQ) - nUQ : 0) + col0;
        }
        const bool same = (sub == 0 || sub == 3);
        const double ratio = sub == 0 ? ratio_uv : ratio_vu;

        double sol[2][TP][MT_Q], in[2][TP][MT_Q];
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int r = 0; r < TP; r++)
#pragma unroll
                for (int q = 0; q < MT_Q; q++) { sol[h][r][q] = 0.0; in[h][r][q] = 0.0; }

        for (uint32_t pt0 = 0; pt0 < npts; pt0 += chunk) {
            const uint32_t cn = min(chunk, npts - pt0);
            if (!(single_chunk && round0 > 0)) {   // a class whose slabs cover all points is staged once for all rounds
                __syncthreads();
                // ---- stage the slabs: curl_f and val_f of every function at the chunk's points (the sampler applied per block).
                // One task = (function column, quadrature row m); the n loop runs inside so N_i(m) / T_i(m) are loaded once.
                const uint32_t m_lo = pt0 / nv, m_hi = (pt0 + cn - 1) / nv;
                for (int side = 0; side < (c.local ? 1 : 2); side++) {
                    const uint32_t stride = side ? strideQ : strideP, nF = side ? nQ : nP, nUF = side ? nUQ : nUP;
                    const uint32_t loff = side ? LQ.off : LP.off;
                    const double* tu = side ? tQu : tPu; const double* tv = side ? tQv : tPv;
                    const double jiu = side ? jiuQ : jiuP, jiv = side ? jivQ : jivP;
                    // derivative scale = the OTHER function's para_scale (integrals.rs:44,47); Q's is (1,1), P's is (su,sv)
                    const double ps0 = side ? c.su : 1.0, ps1 = side ? c.sv : 1.0;
                    double* sC = side ? s_CQ : s_CP; double* sF = side ? s_FQ : s_FP;
                    const uint32_t padU = pad4(nUF);
                    for (int grp = 0; grp < 2; grp++) {
                    const uint32_t c_lo = it.stage[side][grp][0], c_w = it.stage[side][grp][1] - c_lo;   // only the functions this item's tiles touch
                    const uint32_t ntask = (m_hi - m_lo + 1) * c_w;
                    for (uint32_t t = threadIdx.x; t < ntask; t += blockDim.x) {
                        const uint32_t mi = t / c_w, col = c_lo + (t - mi * c_w), m = m_lo + mi;
                        const uint32_t n_lo = (m == m_lo) ? pt0 - m_lo * nv : 0u;
                        const uint32_t n_hi = (m == m_hi) ? pt0 + cn - 1 - m_hi * nv : nv - 1;
                        double* dC = sC + (size_t)(m * nv + n_lo - pt0) * stride + col;
                        double* dF = sF + (size_t)(m * nv + n_lo - pt0) * stride + col;
                        if (col < nUF) {
                            const uint32_t i = g.spec_i[loff + col], j = g.spec_j[loff + col];
                            const double Ni = tu[(0 * g.NO + i) * g.NPT + m];
                            const double jN = jiu * Ni;
                            const double* Tj = tv + 2 * AS + j * g.NPT; const double* Tdj = tv + 3 * AS + j * g.NPT;
                            for (uint32_t n = n_lo; n <= n_hi; n++, dC += stride, dF += stride) {
                                *dC = -((jiu * (Ni * Tdj[n])) * ps0);
                                *dF = jN * Tj[n];
                            }
                        } else if (col >= padU && col - padU < nF - nUF) {
                            const uint32_t a = nUF + (col - padU);
                            const uint32_t i = g.spec_i[loff + a], j = g.spec_j[loff + a];
                            const double Ti = tu[2 * AS + i * g.NPT + m], Tdi = tu[3 * AS + i * g.NPT + m];
                            const double jT = jiv * Ti;
                            const double* Nj = tv + (0 * g.NO + j) * g.NPT;
                            for (uint32_t n = n_lo; n <= n_hi; n++, dC += stride, dF += stride) {
                                *dC = (jiv * (Tdi * Nj[n])) * ps1;
                                *dF = jT * Nj[n];
                            }
                        } else {
                            for (uint32_t n = n_lo; n <= n_hi; n++, dC += stride, dF += stride) { *dC = 0.0; *dF = 0.0; }
                        }
                    }
                    }
                }
                __syncthreads();
            }
            if (active) {
                uint32_t m = pt0 / nv, n = pt0 - m * nv, pl = 0;
                const double* cp = s_CP + prow; const double* fp = s_FP + prow;
                const double* cq = s_CQ + pcol; const double* fq = s_FQ + pcol;
                while (pl < cn) {
                    const uint32_t run = min(cn - pl, nv - n);
                    if (same) contract_run<true, TP>(cp, cq, fp, fq, strideP, strideQ, s_vw + n, run, ratio, maxdet, in);
                    else contract_run<false, TP>(cp, cq, fp, fq, strideP, strideQ, s_vw + n, run, ratio, maxdet, in);
                    pl += run; n += run;
                    if (n == nv) {   // end of the inner (v) loop: solution += inner_solution * u_w (glq.rs:29)
                        const double uw = s_uw[m];
#pragma unroll
                        for (int h = 0; h < 2; h++)
#pragma unroll
                            for (int r = 0; r < TP; r++)
#pragma unroll
                                for (int q = 0; q < MT_Q; q++) { sol[h][r][q] = sol[h][r][q] + in[h][r][q] * uw; in[h][r][q] = 0.0; }
                        n = 0; m++;
                    }
                }
            }
        }
        if (active) {
            if (same) {
#pragma unroll
                for (int r = 0; r < TP; r++) {
                    const uint32_t a = row0 + r;
                    if (a >= row_end) continue;
#pragma unroll
                    for (int q = 0; q < MT_Q; q++) {
                        const uint32_t b = col0 + q;
                        if (b >= col_end) continue;
                        out[(size_t)a * nQ + b] = make_double2(coefA * sol[0][r][q], coefB * sol[1][r][q]);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < TP; r++) {
                    const uint32_t a = row0 + r;
                    if (a >= row_end) continue;
#pragma unroll
                    for (int q = 0; q < MT_QX; q++) {
                        const uint32_t b = col0 + q;
                        if (b >= col_end) continue;
                        // cross-direction mass entries: every integrand term is a signed zero, the quadrature returns +0.0 (integrals.rs:318-339)
                        out[(size_t)a * nQ + b] = make_double2(coefA * sol[q >> 1][r][q & 1], coefB * 0.0);
                    }
                }
            }
        }
    }
    if (!g.follows_sampler) cudaGridDependencySynchronize();   // do not complete before the first integrator grid has
}


// ---------------------------------------------------------------------------------------------------- K2, warp-specialised
// Persistent form of the integrator for the throughput shape (4-row tiles): K2_WS_CONS_WARPS contraction warps and one staging warp
// per CTA, two CTAs per SM.  The staging warp takes work items from a global counter, computes the item's class constants, and fills
// a ring of K2_WS_NBUF slab buffers chunk by chunk; the contraction warps consume the chunks in the same order.  Hand-over is by
// mbarriers (full[b]: one arrival by the staging warp; empty[b]: one arrival per contraction warp), so the contraction warps never
// stage and never meet a CTA-wide barrier: while they contract chunk k the staging warp prepares chunk k+1 -- of the same round, of
// the next round, or of the next work item (its descriptor loads, FP64 quotients and first slab included).  Same tiles, same
// per-pair operation order as k2_exact_kernel: results are bit-identical.
// One work unit of the persistent kernel is a PACK of up to K2_PACK_MAX work items (segments) that share a round: small items of
// different classes whose micro-tiles together fill the contraction threads (plan_types.h PackDesc).  A big item is a pack of one.
struct alignas(16) WsSeg {
    WorkItem it;
    SubBlocks sb;
    double ratio_uv, ratio_vu, maxdet, coefA, coefB;
    double jiuP, jivP, jiuQ, jivQ, su, sv;       // staging: jac_inv entries and P's para_scale
    unsigned long long v_off;
    uint32_t nP, nUP, nQ, nUQ, strideP, strideQ, local;
    uint32_t slab_off;                           // doubles per chunk POINT in front of this segment's slabs
    uint32_t same_off, cross_off;                // first thread slot of the segment inside the same- / cross-direction group of the round
    uint32_t listP_off, listQ_off, tabPu, tabPv; // staging: order lists and P-side tables
};
struct alignas(16) WsCtx {
    WsSeg seg[K2_PACK_MAX];
    uint32_t n_seg, chunk_rows, n_slots, gap, n_same, end;   // n_same: same-direction tiles of all segments (they come first)
    uint32_t n_cols, tab_u, tab_v, pad0, pad1, pad2;         // staging: entries of the pack's column table, scaled P-side tables (0xffffffff: none)
};
// pack contexts in flight: the staging warps run at most NBUF chunks (>= packs) ahead of a contraction warp that may still be in the
// epilogue of the pack before those, and prepare one more pack in their slack time
constexpr int K2_WS_NCTX = K2_WS_NBUF + 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {   // release at CTA scope: the arriving thread's prior writes are visible to whoever observes the phase flip
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the warp up to the hint; a failed probe backs off with nanosleep so that waiting warps leave the issue slots of
    // their scheduler to the warps that stage or contract (a bare retry loop showed up as 2/3 of all executed instructions in ncu)
    uint32_t ok;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
        if (ok) return;
        __nanosleep(64);
    }
}

// One function column of a staging chunk, NB points x NR quadrature rows per step (see ws_stage_chunk).  Straight-line code: a step
// that reaches past the last point or row clamps its indices for loads AND stores, i.e. it recomputes the last point / row and stores
// the identical value again, so the body carries no guards.
template <int NB, int NR>
__device__ __forceinline__ void ws_stage_column(const double* __restrict__ colA, const double* __restrict__ colB, const double* __restrict__ rowA,
                                                const double* __restrict__ rowB, double jj, double ps, bool scaled, int flip, double* __restrict__ sC,
                                                double* __restrict__ sF, uint32_t stride, uint32_t nv, uint32_t nrow) {
    for (uint32_t nb = 0; nb < nv; nb += NB) {
        double ca[NB], cb[NB];
        uint32_t noff[NB];   // slab offset of point n of a row
#pragma unroll
        for (int q = 0; q < NB; q++) { const uint32_t n = min(nb + q, nv - 1); ca[q] = colA[n]; cb[q] = colB[n]; noff[q] = n * stride; }
        for (uint32_t r = 0; r < nrow; r += NR) {
            double fa[NR], fb[NR], t[NR][NB], f[NR][NB];
            uint32_t roff[NR];
#pragma unroll
            for (int k = 0; k < NR; k++) { const uint32_t rk = min(r + k, nrow - 1); fa[k] = rowA[rk]; fb[k] = jj * rowB[rk]; roff[k] = rk * nv * stride; }
#pragma unroll
            for (int k = 0; k < NR; k++)
#pragma unroll
                for (int q = 0; q < NB; q++) t[k][q] = fa[k] * ca[q];
#pragma unroll
            for (int k = 0; k < NR; k++)
#pragma unroll
                for (int q = 0; q < NB; q++) t[k][q] = jj * t[k][q];
            if (scaled) {
#pragma unroll
                for (int k = 0; k < NR; k++)
#pragma unroll
                    for (int q = 0; q < NB; q++) t[k][q] = t[k][q] * ps;
            }
#pragma unroll
            for (int k = 0; k < NR; k++)
#pragma unroll
                for (int q = 0; q < NB; q++) f[k][q] = fb[k] * cb[q];
#pragma unroll
            for (int k = 0; k < NR; k++)
#pragma unroll
                for (int q = 0; q < NB; q++) {
                    sC[roff[k] + noff[q]] = __hiloint2double(__double2hiint(t[k][q]) ^ flip, __double2loint(t[k][q]));
                    sF[roff[k] + noff[q]] = f[k][q];
                }
        }
    }
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking probe of a phase
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// One staging task: a slab column (one function of one segment and side) with its table rows and constants.
struct WsTask {
    const double* colA; const double* colB; const double* rowA; const double* rowB;   // n-axis / m-axis factors of curl and value
    double jj, ps;
    double* sC; double* sF;     // the column's first entry in the curl / value slab of this chunk
    uint32_t stride;
    int flip;
    bool scaled;
};

__device__ __forceinline__ WsTask ws_task(const K2Args& g, const WsCtx& c, uint32_t e, const double* s_tab, double* buf, uint32_t chunk, uint32_t m0) {
    const uint32_t TS = ws_tab_stride(g.NPT), AS = g.NO * TS;
    const uint32_t col = e & 0xffffu, side = (e >> 16) & 1u;
    const WsSeg& sg = c.seg[(e >> 17) & 3u];
    const uint32_t strideP = sg.strideP, strideQ = sg.strideQ;
    const uint32_t nUF = side ? sg.nUQ : sg.nUP;
    WsTask k;
    k.stride = side ? strideQ : strideP;
    double* seg_buf = buf + (size_t)chunk * sg.slab_off;
    // slabs of a segment: C_P [chunk][strideP], F_P, then (non-local) C_Q [chunk][strideQ], F_Q
    k.sC = seg_buf + (side ? (size_t)chunk * 2 * strideP : 0) + col;
    k.sF = k.sC + (size_t)chunk * k.stride;
    // table cache slots: 0 / 1 = the scaled u / v tables of a non-local P side, 2 / 3 = the unscaled tables (every Q side, local P sides)
    const uint32_t slot_u = (side || sg.local) ? 2u : 0u;
    const double* tu = s_tab + (size_t)slot_u * 3 * AS;          // cached tables hold N, T, T' (N' is never sampled on this path)
    const double* tv = s_tab + (size_t)(slot_u + 1) * 3 * AS;
    const bool isU = col < pad4(nUF);
    const uint32_t i = (e >> 20) & 31u, j = (e >> 25) & 31u;   // the column's (i, j) orders travel in its table entry (pack set-up)
    // U-directed: curl = -((jiu * (N_i(m) * T'_j(n))) * ps0), val = (jiu * N_i(m)) * T_j(n)        (basis.rs:225-242)
    // V-directed: curl =   (jiv * (T'_i(m) * N_j(n))) * ps1,  val = (jiv * T_i(m)) * N_j(n)        (basis.rs:230-252)
    // cache layout: [N | T | T'] x order x point
    k.rowA = tu + (isU ? 0 : 2) * AS + i * TS + m0;   // factor of the curl taken at m: N_i | T'_i
    k.rowB = tu + (isU ? 0 : 1) * AS + i * TS + m0;   // factor of the value taken at m: N_i | T_i
    k.colA = tv + (isU ? 2 : 0) * AS + j * TS;        // factor of the curl taken at n: T'_j | N_j
    k.colB = tv + (isU ? 1 : 0) * AS + j * TS;        // factor of the value taken at n: T_j | N_j
    k.jj = isU ? (side ? sg.jiuQ : sg.jiuP) : (side ? sg.jivQ : sg.jivP);
    // derivative scale = the OTHER function's para_scale (integrals.rs:44,47); Q's is (1,1), P's is (su,sv)
    k.ps = side ? (isU ? sg.su : sg.sv) : 1.0;
    k.scaled = k.ps != 1.0;                    // x * 1.0 == x bit for bit: the product is skipped
    k.flip = isU ? (int)0x80000000 : 0;        // IEEE negation = sign-bit flip, done in the integer pipe
    return k;
}

// Stages the slab columns a pack needs for the quadrature rows [m0, m0 + nrow) with the lanes of the staging warps (same values, same
// operation order as the staging pass of k2_exact_kernel).  One task = one function column of one segment for the whole chunk, taken
// from the pack's column table (column | side | segment | i | j), so the lanes stay busy whatever the widths of the segments: the
// v-axis table values of four points are held in registers while the rows m run inside, and the inner body is four FP64 operations and
// two shared-memory stores per (function, point) with no loads in the dependent chain.
template <int PROD>
__device__ __forceinline__ void ws_stage_chunk(const K2Args& g, const WsCtx& c, const uint32_t* s_cols, uint32_t n_cols, const double* s_tab,
                                               double* buf, uint32_t chunk, uint32_t m0, uint32_t nrow, uint32_t lane) {
    const uint32_t nv = g.nv;
    const uint32_t npair = nrow & ~1u;
    // eight independent chains per step, written stage by stage so that the in-order issue of a staging warp never waits on the
    // 8-cycle FP64 latency: 4 points x 2 quadrature rows, then 4 points x the odd last row
    for (uint32_t t = lane; t < n_cols; t += PROD * 32) {   // `lane`: index across the staging warps
        const WsTask a = ws_task(g, c, s_cols[t], s_tab, buf, chunk, m0);
        if (npair) ws_stage_column<4, 2>(a.colA, a.colB, a.rowA, a.rowB, a.jj, a.ps, a.scaled, a.flip, a.sC, a.sF, a.stride, nv, npair);
        if (nrow & 1u) ws_stage_column<4, 1>(a.colA, a.colB, a.rowA + npair, a.rowB + npair, a.jj, a.ps, a.scaled, a.flip, a.sC + (size_t)npair * nv * a.stride,
                                             a.sF + (size_t)npair * nv * a.stride, a.stride, nv, 1u);
    }
}

// Copies the N, T, T' arrays of sampled table `id` (K1 layout [N | N' | T | T'] x order x point) into a cache slot.
__device__ __forceinline__ void ws_cache_table(const K2Args& g, uint32_t id, double* dst, uint32_t lane) {
    const uint32_t AS = g.NO * g.NPT, TS = ws_tab_stride(g.NPT), AC = g.NO * TS;
    const double* src = g.tabs + (size_t)id * 4 * AS;
#pragma unroll 4
    for (uint32_t a = lane; a < AS; a += 32) {
        const uint32_t c = a + (a / g.NPT) * (TS - g.NPT);   // row o = a / NPT starts at o * TS
        dst[c] = src[a]; dst[AC + c] = src[2 * AS + a]; dst[2 * AC + c] = src[3 * AS + a];
    }
}

// One pass of a TP x 2 sub-tile over one quadrature row: inner += ((p * q) [* scale]) * v_w[n] for n = 0 .. nv-1 (glq.rs:24-28), then
// solution += inner * u_w[m] (glq.rs:29).  A same-direction tile runs it twice per row (curl slabs with the uv / vu ratio for A, value
// slabs with max(det) for B), a cross-direction tile twice on the curl slabs (columns 0-1, columns 2-3; no ratio factor,
// integrals.rs:53-84).  Running the passes one after the other keeps a single set of inner accumulators live, which leaves the
// scheduler the registers to interleave all eight chains of a pass and to fetch the next points' operands early.
// MODE 2 multiplies the WEIGHT by `scale` instead of every product: bit-identical when `scale` is a power of two, because then
// both x * scale and scale * w are exact and ((x * scale) * w) = fl(x * scale * w) = (x * (scale * w)) (no overflow / underflow: the
// exponent of scale is within +-64, plan_types.h is_pow2_scale) -- one DMUL per point instead of eight.  Ratios of dyadically refined Elems are powers
// of two on every mesh, max(det) is one when the base Elements' sides are (the reference's test meshes a and b; not c: 4.2 x 4.2).
template <int MODE, int TP>   // 0: no scale (cross-direction), 1: scale every product, 2: scale the weight (power-of-two scale)
__device__ __forceinline__ void ws_row_pass(const double* __restrict__ p, const double* __restrict__ q, uint32_t strideP, uint32_t strideQ,
                                            const double* __restrict__ vw, uint32_t nv, double scale, double uw, double (&sol)[TP][MT_Q]) {
    double in[TP][MT_Q];
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) in[r][c] = 0.0;
#pragma unroll 4
    for (uint32_t n = 0; n < nv; n++) {
        double pv[TP], qv[MT_Q], t[TP][MT_Q];
        load_rows<TP>(p, pv); load_rows<MT_Q>(q, qv);
        const double w = MODE == 2 ? vw[n] * scale : vw[n];
#pragma unroll
        for (int r = 0; r < TP; r++)
#pragma unroll
            for (int c = 0; c < MT_Q; c++) t[r][c] = pv[r] * qv[c];
        if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < TP; r++)
#pragma unroll
                for (int c = 0; c < MT_Q; c++) t[r][c] = t[r][c] * scale;
        }
#pragma unroll
        for (int r = 0; r < TP; r++)
#pragma unroll
            for (int c = 0; c < MT_Q; c++) t[r][c] = t[r][c] * w;
#pragma unroll
        for (int r = 0; r < TP; r++)
#pragma unroll
            for (int c = 0; c < MT_Q; c++) in[r][c] = in[r][c] + t[r][c];
        p += strideP; q += strideQ;
    }
#pragma unroll
    for (int r = 0; r < TP; r++)
#pragma unroll
        for (int c = 0; c < MT_Q; c++) sol[r][c] = sol[r][c] + in[r][c] * uw;
}

#ifdef FEM2D_WS_PROFILE
// tuning build only: cycle counts summed over CTAs. 0 staging warp: item setup, 1 waiting for an empty buffer, 2 staging;
// 3 contraction warp 0: waiting for a full buffer, 4 contracting (+ epilogue, decode); 5 items; 6 chunks; 7 first chunk of a pack;
// 8-12 pack set-up: claim + syncs, descriptors -> context, offsets + table cache, column table, (i, j) orders
__device__ unsigned long long g_ws_prof[16];
#define WS_T(var) const long long var = clock64()
#define WS_ADD(k, v) do { if (lane == 0) atomicAdd(&g_ws_prof[k], (unsigned long long)(v)); } while (0)
#else
#define WS_T(var)
#define WS_ADD(k, v)
#endif

// Per-class constants, once per plan.  HierCurlBasisFn::defined_over, basis.rs:395-413; M2D::det / inverse, space.rs:138-147; the same
// expressions (and the same separately rounded operations) as the prologue of k2_exact_kernel.
__global__ void class_geom_kernel(const ClassDesc* __restrict__ classes, uint32_t n, ClassGeom* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const ClassDesc& cd = classes[c];
    const double dxP = cd.dxP, dyP = cd.dyP, dxQ = cd.dxQ, dyQ = cd.dyQ;
    const double detP = dxP * dyP - 0.0 * 0.0, detQ = dxQ * dyQ - 0.0 * 0.0;
    const double ge = (double)(detP >= detQ), lt = (double)(detP < detQ);
    ClassGeom o;
    o.jiuP = dyP / detP; o.jivP = dxP / detP;      // jac_inv.u[0] = dy_dv / det, jac_inv.v[1] = dx_du / det
    o.jiuQ = dyQ / detQ; o.jivQ = dxQ / detQ;
    o.ratio_uv = ge * (dxP / dyP) + lt * (dxQ / dyQ);   // max_uv_ratios integrals.rs:250-259, basis.rs:341-343
    o.ratio_vu = ge * (dyP / dxP) + lt * (dyQ / dxQ);   // max_vu_ratios integrals.rs:262-271, basis.rs:346-348
    o.maxdet = detP > detQ ? detP : detQ;               // partial_max integrals.rs:421-423
    o.coefA = 1.0 / cd.mu;                              // integrals.rs:37
    o.coefB = cd.eps * (cd.su * cd.sv) * (1.0 * 1.0);   // eps * p.glq_scale() * q.glq_scale() integrals.rs:303-305
    out[c] = o;
}

// Pack set-up by the first staging warp, in two steps of about one staged chunk's time each: (1) work items, classes and per-class
// constants -> context `c` (one lane per segment), thread-slot and slab offsets; (2) the column table with the (i, j) orders of every
// staged function.  Functions of their own: each runs once per pack from two places -- ahead of time, when the staging warps would
// otherwise wait for a free ring buffer, or at the start of the pack.
template <int TP>
__device__ __noinline__ void ws_setup_context(const PackDesc* __restrict__ packs, uint32_t n_packs, uint32_t idx, const WorkItem* __restrict__ items,
                                           const ClassDesc* __restrict__ classes, const ClassGeom* __restrict__ geom,
                                           WsCtx* cp, uint32_t nu, uint32_t nv, uint32_t buf_doubles, uint32_t lane) {
    WsCtx& c = *cp;
    if (idx >= n_packs) {   // end marker: travels through the ring like a chunk
        if (lane == 0) c.end = 1u;
        __syncwarp();
        return;
    }
    WS_T(t_s0);
    const PackDesc pk = packs[idx];
    // ---- one lane per segment: work item, class, per-class constants (ClassGeom) -> context
    if (lane < pk.n) {
        WsSeg& sg = c.seg[lane];
        const WorkItem it = items[pk.first + lane];
        const ClassDesc cd = classes[it.cls];
        const ClassGeom cg = geom[it.cls];
        const uint32_t nP = cd.lp.n, nUP = cd.lp.nU, nQ = cd.lq.n, nUQ = cd.lq.nU;
        sg.it = it;
        sg.sb = make_subblocks(nP, nUP, nQ, nUQ, cd.local, TP);
        sg.jiuP = cg.jiuP; sg.jivP = cg.jivP; sg.jiuQ = cg.jiuQ; sg.jivQ = cg.jivQ;
        sg.ratio_uv = cg.ratio_uv; sg.ratio_vu = cg.ratio_vu; sg.maxdet = cg.maxdet; sg.coefA = cg.coefA; sg.coefB = cg.coefB;
        sg.su = cd.su; sg.sv = cd.sv;
        sg.v_off = cd.v_off;
        sg.nP = nP; sg.nUP = nUP; sg.nQ = nQ; sg.nUQ = nUQ; sg.local = cd.local;
        sg.strideP = pad4(nUP) + pad4(nP - nUP);
        sg.strideQ = cd.local ? sg.strideP : pad4(nUQ) + pad4(nQ - nUQ);
        sg.listP_off = cd.lp.off; sg.listQ_off = cd.lq.off; sg.tabPu = cd.tabPu; sg.tabPv = cd.tabPv;
    }
    __syncwarp();
    WS_T(t_s1); WS_ADD(9, t_s1 - t_s0);
    // ---- slot / slab offsets of the segments, chunk size
    uint32_t tot_stride = 0, n_same = 0, n_cross = 0, tab_u = 0xffffffffu, tab_v = 0xffffffffu;
    for (uint32_t s2 = 0; s2 < pk.n; s2++) {
        WsSeg& sg = c.seg[s2];
        if (lane == 0) { sg.slab_off = 2 * tot_stride; sg.same_off = n_same; sg.cross_off = n_cross; }
        tot_stride += sg.strideP + (sg.local ? 0u : sg.strideQ);
        n_same += sg.it.n_same; n_cross += sg.it.mt_count - sg.it.n_same;
        if (!sg.local) { tab_u = sg.tabPu; tab_v = sg.tabPv; }   // the planner packs non-local segments of one table pair only
    }
    // whole quadrature rows per chunk (the launch makes sure one row of the widest pack fits a ring buffer)
    const uint32_t chunk_rows = min(nu, buf_doubles / (2 * nv * tot_stride));
    const uint32_t gap = item_gap(n_same, n_same + n_cross), n_slots = n_same + n_cross + gap;
    if (lane == 0) { c.n_seg = pk.n; c.chunk_rows = chunk_rows; c.n_slots = n_slots; c.gap = gap; c.n_same = n_same; c.end = 0u; c.tab_u = tab_u; c.tab_v = tab_v; }
    WS_T(t_s2); WS_ADD(10, t_s2 - t_s1);
    __syncwarp();
}

__device__ __noinline__ void ws_setup_columns(const uint8_t* __restrict__ spec_i, const uint8_t* __restrict__ spec_j, WsCtx* cp, uint32_t* s_cols, uint32_t lane) {
    WsCtx& c = *cp;
    if (c.end) return;
    const uint32_t n_seg = c.n_seg;
    WS_T(t_s2);
    // ---- column table: every slab column the pack's tiles touch, as column | side << 16 | segment << 17 | i << 20 | j << 25
    uint32_t n_cols = 0;
    for (uint32_t s2 = 0; s2 < n_seg; s2++) {
        const WsSeg& sg = c.seg[s2];
        for (uint32_t side = 0; side < (sg.local ? 1u : 2u); side++)
            for (uint32_t grp = 0; grp < 2; grp++) {
                const uint32_t c_lo = sg.it.stage[side][grp][0], c_w = sg.it.stage[side][grp][1] - c_lo;   // only the functions this item's tiles touch
                for (uint32_t k = lane; k < c_w; k += 32) s_cols[n_cols + k] = s2 << 17 | side << 16 | (c_lo + k);
                n_cols += c_w;
            }
    }
    __syncwarp();
    WS_T(t_s3); WS_ADD(11, t_s3 - t_s2);
    // (i, j) orders of the columns' functions, straight from the plan's pooled BasisSpec lists: 8 columns per lane and batch, so
    // that all global loads of a batch are in flight together (one L2 round trip per 256 columns)
    for (uint32_t t0 = 0; t0 < n_cols; t0 += 8 * 32) {
        uint32_t e[8], vi[8], vj[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t t = t0 + u * 32 + lane;
            e[u] = 0; vi[u] = 0; vj[u] = 0;
            if (t < n_cols) {
                e[u] = s_cols[t];
                const WsSeg& sg = c.seg[(e[u] >> 17) & 3u];
                const uint32_t side = (e[u] >> 16) & 1u, col = e[u] & 0xffffu;
                const uint32_t nF = side ? sg.nQ : sg.nP, nUF = side ? sg.nUQ : sg.nUP, padU = pad4(nUF);
                // a padding column of a 4-wide tile row takes the values of the group's last function: only pairs beyond the block's
                // last row / column read it, and their results are never stored
                const uint32_t a = col < padU ? min(col, nUF - 1) : nUF + min(col - padU, nF - nUF - 1);
                const uint32_t off = (side ? sg.listQ_off : sg.listP_off) + a;
                vi[u] = spec_i[off]; vj[u] = spec_j[off];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t t = t0 + u * 32 + lane;
            if (t < n_cols) s_cols[t] = e[u] | vi[u] << 20 | vj[u] << 25;
        }
    }
    WS_T(t_s4); WS_ADD(12, t_s4 - t_s3);
    if (lane == 0) c.n_cols = n_cols;
    __syncwarp();
}

constexpr uint32_t K2_WS_SM_SLOTS = 1024;
__device__ uint32_t g_ws_sm_arrivals[K2_WS_SM_SLOTS];   // CTAs of the persistent integrator that have started on each SM, ever (only the parity is used)

// FOLD bit 0: the uv / vu ratios of every class of the plan are powers of two, bit 1: so is every max(det) -- decided per plan on the
// host (HostPlan::ws_fold); such a scale multiplies the quadrature weight instead of every product (ws_row_pass), bit for bit the same.
template <int TP, int PROD, int FOLD>
__global__ void __maxnreg__(K2_WS_MAXREG) k2_ws_kernel(const K2Args g, const PackDesc* __restrict__ packs, const uint32_t n_packs, uint32_t* __restrict__ work_counter) {
    extern __shared__ __align__(16) double smem[];
    constexpr int K2_WS_PROD_WARPS = PROD, K2_WS_CONS_WARPS = K2_WS_WARPS - PROD;   // staging / contraction warps of this instantiation
    constexpr uint32_t CONS = K2_WS_CONS_WARPS * 32;
    double* s_uw = smem;                       // [128]
    double* s_vw = smem + 128;                 // [128]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + 256);   // [NBUF]
    uint64_t* s_empty = s_full + K2_WS_NBUF;                      // [NBUF]
    WsCtx* s_ctx = reinterpret_cast<WsCtx*>(smem + 256 + 2 * K2_WS_NBUF);
    static_assert(sizeof(WsCtx) % 16 == 0, "context ring keeps the slabs 16-byte aligned");
    // staging-warp caches: four sampled tables (slots 0 / 1: scaled u / v tables of the non-local P sides of the current pack; 2 / 3: the
    // unscaled u / v tables) and the column table of the current pack
    const uint32_t AS3 = 3 * g.NO * ws_tab_stride(g.NPT);
    double* s_tab = smem + 256 + 2 * K2_WS_NBUF + K2_WS_NCTX * sizeof(WsCtx) / sizeof(double);
    uint32_t* s_cols = reinterpret_cast<uint32_t*>(s_tab + 4 * (size_t)AS3);   // column table of the current pack
    double* s_slab = reinterpret_cast<double*>(s_cols + 2 * g.col_cap);   // two column tables: the current pack's and the next one's
    const uint32_t buf_doubles = g.slab_doubles;   // per ring buffer
    const uint32_t lane = threadIdx.x % 32;
    const uint32_t nu = g.nu, nv = g.nv;
    // Roles by SM sub-partition.  A warp's scheduler and FP64 pipe are those of its hardware warp slot (%warpid % 4, measured:
    // scripts/micro/smsp_map.cu), a CTA of eight warps has two warps on each of the four sub-partitions, and which warp index sits
    // where differs between the two CTAs of an SM.  With the staging warps taken by index the contraction warps of an SM spread
    // 2 / 3 / 4 / 3 over the sub-partitions (two staging warps per CTA) and the fullest one paces every chunk.  So the first CTA on
    // an SM stages on sub-partitions 0 (and 1), the second on 2 (and 3): 3 contraction warps + 1 staging warp everywhere with two
    // staging warps per CTA.  Only the warp -> role map changes; tiles, threads slots and arithmetic do not.
    __shared__ uint32_t s_smsp[K2_WS_WARPS], s_second;
    {
        uint32_t hw;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
        if (lane == 0) s_smsp[threadIdx.x / 32] = hw & 3u;
    }

    if (!g.follows_sampler) cudaTriggerProgrammaticLaunchCompletion();
    if (threadIdx.x == 0) {
        for (int b = 0; b < K2_WS_NBUF; b++) { mbar_init(&s_full[b], K2_WS_PROD_WARPS * 32); mbar_init(&s_empty[b], K2_WS_CONS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_second = atomicAdd(&g_ws_sm_arrivals[smid % K2_WS_SM_SLOTS], 1u) & 1u;   // two consecutive arrivals on an SM always differ
    }
    for (uint32_t k = threadIdx.x; k < nu; k += blockDim.x) s_uw[k] = g.glq[128 + k];
    for (uint32_t k = threadIdx.x; k < nv; k += blockDim.x) s_vw[k] = g.glq[384 + k];
    if (g.follows_sampler) {
        cudaGridDependencySynchronize();   // the sampler's tables (and, transitively, the previous call's readers of V) are complete
        cudaTriggerProgrammaticLaunchCompletion();
    }
    __syncthreads();

    // logical warp index: staging warps first (0 .. PROD-1), then the contraction warps in index order
    uint32_t warp;
    {
        const uint32_t me = threadIdx.x / 32;
        uint32_t taken = 0;   // bit w: warp w stages
        int mine = -1;
#pragma unroll
        for (int k = 0; k < K2_WS_PROD_WARPS; k++) {
            const uint32_t want = (2u * s_second + (uint32_t)k) & 3u;
            uint32_t pick = 0xffffffffu;
            for (uint32_t w = 0; w < (uint32_t)K2_WS_WARPS; w++)
                if (g.roles_by_subpartition && pick == 0xffffffffu && !(taken >> w & 1u) && s_smsp[w] == want) pick = w;
            for (uint32_t w = 0; w < (uint32_t)K2_WS_WARPS; w++)   // no warp there (never seen): lowest free index
                if (pick == 0xffffffffu && !(taken >> w & 1u)) pick = w;
            taken |= 1u << pick;
            if (pick == me) mine = k;
        }
        warp = mine >= 0 ? (uint32_t)mine : (uint32_t)K2_WS_PROD_WARPS + me - (uint32_t)__popc(taken & ((1u << me) - 1u));
    }

    uint32_t stage = 0, phase = 0, ci = 0;
    if (warp < K2_WS_PROD_WARPS) {
        // ================================================================================================ staging warps
        // (the first warps of the CTA: the warp schedulers favour the oldest warps, and staging must stay ahead of the contraction warps)
        // With more than one staging warp the pack set-up is done by warp 0 and the staging tasks of every chunk are dealt round-robin;
        // the staging warps meet at a named barrier after the set-up and before the next set-up overwrites the caches.
        const uint32_t plane = warp * 32 + lane;   // lane index across the staging warps
        auto prod_sync = [] { if (K2_WS_PROD_WARPS > 1) asm volatile("bar.sync 1, %0;" ::"r"(K2_WS_PROD_WARPS * 32) : "memory"); else __syncwarp(); };
        if (warp == 0) { ws_cache_table(g, 0u, s_tab + 2 * (size_t)AS3, lane); ws_cache_table(g, 1u, s_tab + 3 * (size_t)AS3, lane); }   // tables 0 / 1: unscaled u / v points
        uint32_t cached_u = 0xffffffffu, cached_v = 0xffffffffu;                                 // table ids held by slots 0 / 1
        // Packs are claimed one ahead (the atomic's round trip runs behind the staging of the current pack), and the claimed pack is set
        // up ahead of time when the staging warps find their next ring buffer still in use: its context and column table go to the next
        // slots of their rings, so the contraction warps rarely wait for the first chunk of a pack.
        uint32_t cur_idx = 0, next_idx = 0, cb = 0;   // cb: column-table buffer of the current pack
        uint32_t cur_ready = 0, next_ready = 0;   // set-up steps done: 0 none, 1 context, 2 context + column table
        if (warp == 0) { if (lane == 0) cur_idx = atomicAdd(work_counter, 1u); cur_idx = __shfl_sync(0xffffffffu, cur_idx, 0); }
        for (;;) {
            WS_T(t_item0);
            prod_sync();   // every staging warp is done with the previous pack's column table and table cache
            WsCtx& c = s_ctx[ci];
            const uint32_t* cols = s_cols + cb * g.col_cap;
            if (warp == 0) {
                if (cur_ready < 1u) ws_setup_context<TP>(packs, n_packs, cur_idx, g.items, g.classes, g.geom, &c, nu, nv, buf_doubles, lane);
                if (cur_ready < 2u) ws_setup_columns(g.spec_i, g.spec_j, &c, s_cols + cb * g.col_cap, lane);
                if (!c.end) {
                    if (lane == 0) next_idx = atomicAdd(work_counter, 1u);   // consumed when the next pack is set up
                    // the scaled P-side tables (the ancestor sampled over the descendant, basis.rs:372-393)
                    const uint32_t tab_u = c.tab_u, tab_v = c.tab_v;
                    if (tab_u != 0xffffffffu && cached_u != tab_u) { ws_cache_table(g, tab_u, s_tab, lane); cached_u = tab_u; }
                    if (tab_v != 0xffffffffu && cached_v != tab_v) { ws_cache_table(g, tab_v, s_tab + AS3, lane); cached_v = tab_v; }
                }
                next_ready = 0;
            }
            prod_sync();
            if (c.end) {
                mbar_wait(&s_empty[stage], phase ^ 1u);
                mbar_arrive(&s_full[stage]);
                break;
            }
            const uint32_t n_cols = c.n_cols, chunk_rows = c.chunk_rows, n_slots = c.n_slots;
            WS_T(t_item1); if (warp == 0) { WS_ADD(0, t_item1 - t_item0); WS_ADD(5, 1); }
            const uint32_t chunk = chunk_rows * nv;
            for (uint32_t round0 = 0; round0 < n_slots; round0 += K2_WS_TPT * CONS) {
                for (uint32_t m0 = 0; m0 < nu; m0 += chunk_rows) {
                    const uint32_t nrow = min(chunk_rows, nu - m0);
                    WS_T(t_w0);
                    if (warp == 0 && next_ready < 2u && !__any_sync(0xffffffffu, mbar_test(&s_empty[stage], phase ^ 1u))) {
                        // slack: the ring buffer is still being read -> one step of the set-up of the pack claimed at the start of this one
                        WsCtx* cn = &s_ctx[ci + 1 == K2_WS_NCTX ? 0 : ci + 1];
                        if (next_ready == 0) ws_setup_context<TP>(packs, n_packs, __shfl_sync(0xffffffffu, next_idx, 0), g.items, g.classes, g.geom, cn, nu, nv, buf_doubles, lane);
                        else ws_setup_columns(g.spec_i, g.spec_j, cn, s_cols + (cb ^ 1u) * g.col_cap, lane);
                        next_ready++;
                        WS_ADD(13, 1);
                    }
                    mbar_wait(&s_empty[stage], phase ^ 1u);     // the contraction warps are done with what this buffer held
                    WS_T(t_w1);
                    double* buf = s_slab + (size_t)stage * buf_doubles;
                    ws_stage_chunk<PROD>(g, c, cols, n_cols, s_tab, buf, chunk, m0, nrow, plane);
                    __syncwarp();
                    WS_T(t_w2); if (warp == 0) { WS_ADD(1, t_w1 - t_w0); WS_ADD(2, t_w2 - t_w1); WS_ADD(6, 1); }
                    mbar_arrive(&s_full[stage]);   // every staging lane releases its own slab stores (the staging warps have slack for the 32 arrivals)
                    stage = stage + 1 == K2_WS_NBUF ? 0u : stage + 1; phase ^= (stage == 0u);
                }
            }
            if (warp == 0) { cur_idx = __shfl_sync(0xffffffffu, next_idx, 0); cur_ready = next_ready; }
            ci = ci + 1 == K2_WS_NCTX ? 0u : ci + 1; cb ^= 1u;
        }
    } else {
        // ============================================================================================ contraction warps
        const uint32_t tid = (warp - K2_WS_PROD_WARPS) * 32 + lane;   // 0 .. CONS-1
        for (;;) {
            WS_T(t_f0);
            mbar_wait(&s_full[stage], phase);   // first chunk of the next pack (or the end marker); its context is complete
            WS_T(t_f1);
            if (warp == K2_WS_PROD_WARPS) WS_ADD(7, t_f1 - t_f0);
            const WsCtx& c = s_ctx[ci];
            if (c.end) break;
            const uint32_t n_slots = c.n_slots, gap = c.gap, n_same = c.n_same, n_seg = c.n_seg, chunk_rows = c.chunk_rows;
            const uint32_t chunk = chunk_rows * nv;
            bool first = true;
            // One round = CONS thread slots.  A thread holds two ASSIGNMENTS, one per pass over a quadrature row:
            //   same-direction slot: both are its own 4 x 2 tile (pass 0 accumulates A from the curl slabs, pass 1 B from the value slabs);
            //   cross-direction slot (A only, 4 x 4 tile = two 4 x 2 halves): the 64 halves of a warp's 32 tiles are dealt so that in each
            //   pass the lanes read CONSECUTIVE column pairs -- lane l takes half l % 2 of tile 16 * pass + l / 2 of its warp.  (A lane that
            //   takes both halves of its own tile reads 16 bytes at a 32-byte lane stride: a two-way bank conflict on every Q-side load,
            //   8 instead of 4 LSU cycles, scripts/micro/lds_bw.cu.)  Which lane computes a pair does not enter its value.
            // asg = sub | segment << 2 | half << 4 | row tile << 5 | column tile << 19, 0xffffffff = none; rc = slab column of the first row | of the first column << 16
            static_assert(K2_WS_TPT == 1, "one micro-tile slot per contraction thread and round");
            for (uint32_t round0 = 0; round0 < n_slots; round0 += CONS) {
                uint32_t asg[2], rc[2];
                const uint32_t my_slot = round0 + tid;
                const bool cross_warp = (my_slot & ~31u) >= n_same + gap;   // cross-direction tiles start on a warp boundary (item_gap)
    #pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t slot = cross_warp ? (my_slot & ~31u) + 16u * h + (lane >> 1) : my_slot;
                    const uint32_t half = cross_warp ? (lane & 1u) : 0u;
                    asg[h] = 0xffffffffu; rc[h] = 0;
                    if (h == 1 && !cross_warp) { asg[1] = asg[0]; rc[1] = rc[0]; continue; }
                    if (slot < n_slots && !(slot >= n_same && slot < n_same + gap)) {
                        // segment: the same-direction tiles of all segments come first (segment by segment), then the cross-direction ones
                        const bool is_same = slot < n_same;
                        const uint32_t k = is_same ? slot : slot - n_same - gap;
                        uint32_t sgi = 0;
                        while (sgi + 1 < n_seg && k >= (is_same ? c.seg[sgi + 1].same_off : c.seg[sgi + 1].cross_off)) sgi++;
                        const WsSeg& sg = c.seg[sgi];
                        uint32_t li = is_same ? k - sg.same_off : sg.it.n_same + (k - sg.cross_off), r = 0, sub, rt, ct;
                        while (li >= sg.it.rcount[r]) { li -= sg.it.rcount[r]; r++; }        // which of the item's tile ranges
                        decode_tile(sg.sb, sg.it.rbegin[r] + li, TP, sub, rt, ct);
                        asg[h] = sub | sgi << 2 | half << 4 | rt << 5 | ct << 19;
                        const uint32_t prow = (sub >= 2 ? pad4(sg.nUP) - sg.nUP : 0) + sg.sb.row0[sub] + rt * TP;               // slab column of the canonical row index
                        const uint32_t pcol = ((sub & 1) ? pad4(sg.nUQ) - sg.nUQ : 0) + sg.sb.col0[sub] + ct * mt_width(sub) + half * MT_Q;
                        rc[h] = prow | pcol << 16;
                    }
                }
                // sol[0] / sol[1]: A / B of a same-direction tile; the A values of the two cross-direction halves
                double sol[2][TP][MT_Q];
    #pragma unroll
                for (int h = 0; h < 2; h++)
    #pragma unroll
                    for (int r = 0; r < TP; r++)
    #pragma unroll
                        for (int q = 0; q < MT_Q; q++) sol[h][r][q] = 0.0;
                for (uint32_t m0 = 0; m0 < nu; m0 += chunk_rows) {
                    const uint32_t nrow = min(chunk_rows, nu - m0);
                    WS_T(t_c0);
                    if (!first) mbar_wait(&s_full[stage], phase);
                    first = false;
                    WS_T(t_c1);
                    if (warp == K2_WS_PROD_WARPS) WS_ADD(3, t_c1 - t_c0);
                    const double* buf = s_slab + (size_t)stage * buf_doubles;
                    if (!cross_warp) {
                        if (asg[0] != 0xffffffffu) {
                            const uint32_t sub = asg[0] & 3u;
                            const WsSeg& sg = c.seg[(asg[0] >> 2) & 3u];
                            const uint32_t strideP = sg.strideP, strideQ = sg.strideQ;
                            const double* s_CP = buf + (size_t)chunk * sg.slab_off; const double* s_FP = s_CP + (size_t)chunk * strideP;
                            const double* s_CQ = sg.local ? s_CP : s_FP + (size_t)chunk * strideP;
                            const double* s_FQ = sg.local ? s_FP : s_CQ + (size_t)chunk * strideQ;
                            const double* cp = s_CP + (rc[0] & 0xffffu); const double* cq = s_CQ + (rc[0] >> 16);
                            const double* fp = s_FP + (rc[0] & 0xffffu); const double* fq = s_FQ + (rc[0] >> 16);
                            const double ratio = sub == 0 ? sg.ratio_uv : sg.ratio_vu, maxdet = sg.maxdet;
                            // scales that are powers of two in every class of the plan go into the weights (FOLD, ws_row_pass MODE 2)
                            for (uint32_t r = 0; r < nrow; r++) {
                                const double uw = s_uw[m0 + r];
                                ws_row_pass<(FOLD & 1) ? 2 : 1, TP>(cp, cq, strideP, strideQ, s_vw, nv, ratio, uw, sol[0]);      // A: (curl_p * curl_q) * ratio
                                ws_row_pass<(FOLD & 2) ? 2 : 1, TP>(fp, fq, strideP, strideQ, s_vw, nv, maxdet, uw, sol[1]);     // B: (val_p * val_q) * max(det)
                                cp += (size_t)nv * strideP; cq += (size_t)nv * strideQ; fp += (size_t)nv * strideP; fq += (size_t)nv * strideQ;
                            }
                        }
                    } else {
                        // the two halves may belong to different segments of a pack: each carries its own slabs and strides
                        const double* cp[2]; const double* cq[2]; uint32_t sP[2], sQ[2];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const WsSeg& sg = c.seg[asg[h] == 0xffffffffu ? 0u : (asg[h] >> 2) & 3u];
                            sP[h] = sg.strideP; sQ[h] = sg.strideQ;
                            const double* s_CP = buf + (size_t)chunk * sg.slab_off;
                            const double* s_CQ = sg.local ? s_CP : s_CP + (size_t)chunk * 2 * sP[h];
                            cp[h] = s_CP + (rc[h] & 0xffffu); cq[h] = s_CQ + (rc[h] >> 16);
                        }
                        for (uint32_t r = 0; r < nrow; r++) {
                            const double uw = s_uw[m0 + r];
                            if (asg[0] != 0xffffffffu) ws_row_pass<0, TP>(cp[0], cq[0], sP[0], sQ[0], s_vw, nv, 1.0, uw, sol[0]);
                            if (asg[1] != 0xffffffffu) ws_row_pass<0, TP>(cp[1], cq[1], sP[1], sQ[1], s_vw, nv, 1.0, uw, sol[1]);
                            cp[0] += (size_t)nv * sP[0]; cq[0] += (size_t)nv * sQ[0]; cp[1] += (size_t)nv * sP[1]; cq[1] += (size_t)nv * sQ[1];
                        }
                    }
                    __syncwarp();
                    WS_T(t_c2);
                    if (warp == K2_WS_PROD_WARPS) WS_ADD(4, t_c2 - t_c1);
                    if (lane == 0) mbar_arrive(&s_empty[stage]);   // this warp is done reading the buffer
                    stage = stage + 1 == K2_WS_NBUF ? 0u : stage + 1; phase ^= (stage == 0u);
                }
                if (!cross_warp) {
                    if (asg[0] != 0xffffffffu) {
                        const uint32_t sub = asg[0] & 3u, rt = (asg[0] >> 5) & 0x3fffu, ct = asg[0] >> 19;
                        const WsSeg& sg = c.seg[(asg[0] >> 2) & 3u];
                        const uint32_t row0 = sg.sb.row0[sub] + rt * TP, col0 = sg.sb.col0[sub] + ct * MT_Q;
                        const uint32_t row_end = sg.sb.row0[sub] + sg.sb.rows[sub], col_end = sg.sb.col0[sub] + sg.sb.cols[sub];
                        const uint32_t nQ = sg.nQ;
                        const double coefA = sg.coefA, coefB = sg.coefB;
                        double2* out = g.V + sg.v_off;
    #pragma unroll
                        for (int r = 0; r < TP; r++) {
                            const uint32_t a = row0 + r;
                            if (a >= row_end) continue;
    #pragma unroll
                            for (int q = 0; q < MT_Q; q++) {
                                const uint32_t b = col0 + q;
                                if (b >= col_end) continue;
                                out[(size_t)a * nQ + b] = make_double2(coefA * sol[0][r][q], coefB * sol[1][r][q]);
                            }
                        }
                    }
                } else {
    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (asg[h] == 0xffffffffu) continue;
                        const uint32_t sub = asg[h] & 3u, half = (asg[h] >> 4) & 1u, rt = (asg[h] >> 5) & 0x3fffu, ct = asg[h] >> 19;
                        const WsSeg& sg = c.seg[(asg[h] >> 2) & 3u];
                        const uint32_t row0 = sg.sb.row0[sub] + rt * TP, col0 = sg.sb.col0[sub] + ct * MT_QX + half * MT_Q;
                        const uint32_t row_end = sg.sb.row0[sub] + sg.sb.rows[sub], col_end = sg.sb.col0[sub] + sg.sb.cols[sub];
                        const uint32_t nQ = sg.nQ;
                        const double coefA = sg.coefA, coefB = sg.coefB;
                        double2* out = g.V + sg.v_off;
    #pragma unroll
                        for (int r = 0; r < TP; r++) {
                            const uint32_t a = row0 + r;
                            if (a >= row_end) continue;
    #pragma unroll
                            for (int q = 0; q < MT_Q; q++) {
                                const uint32_t b = col0 + q;
                                if (b >= col_end) continue;
                                // cross-direction mass entries: every integrand term is a signed zero, the quadrature returns +0.0 (integrals.rs:318-339)
                                out[(size_t)a * nQ + b] = make_double2(coefA * sol[h][r][q], coefB * 0.0);
                            }
                        }
                    }
                }
            }
            ci = ci + 1 == K2_WS_NCTX ? 0u : ci + 1;
        }
    }
    if (!g.follows_sampler) cudaGridDependencySynchronize();   // do not complete before the first integrator grid has
}

// ---------------------------------------------------------------------------------------------------------------- FP64 peak
template <int KIND>
__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 1.0000001, b = 1e-9;
    for (int k = 0; k < iters; k++) {
        if (KIND == 0) {
            a0 = __fma_rn(a0, m, b); a1 = __fma_rn(a1, m, b); a2 = __fma_rn(a2, m, b); a3 = __fma_rn(a3, m, b);
            a4 = __fma_rn(a4, m, b); a5 = __fma_rn(a5, m, b); a6 = __fma_rn(a6, m, b); a7 = __fma_rn(a7, m, b);
        } else {
            a0 = __dadd_rn(__dmul_rn(a0, m), b); a1 = __dadd_rn(__dmul_rn(a1, m), b); a2 = __dadd_rn(__dmul_rn(a2, m), b); a3 = __dadd_rn(__dmul_rn(a3, m), b);
            a4 = __dadd_rn(__dmul_rn(a4, m), b); a5 = __dadd_rn(__dmul_rn(a5, m), b); a6 = __dadd_rn(__dmul_rn(a6, m), b); a7 = __dadd_rn(__dmul_rn(a7, m), b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

cudaError_t launch_k1_tables(const Plan& P, int basis_kind, uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st) {
    k1_tables_kernel<<<(unsigned)P.host.tables.size(), 64, 0, st>>>(P.d_tables, P.d_tabs, NO, NPT, P.d_glq, nu, nv, P.host.i_max, P.host.j_max, basis_kind,
                                                                      P.d_work_counter);
    return cudaGetLastError();
}

cudaError_t launch_class_geom(const Plan& P, uint32_t n_classes, cudaStream_t st) {
    if (n_classes == 0) return cudaSuccess;
    class_geom_kernel<<<(n_classes + 127) / 128, 128, 0, st>>>(P.d_classes, n_classes, P.d_class_geom);
    return cudaGetLastError();
}

// One launch of the exact integrator over items [first, first + count) with CTAs of NT threads; max_stride = widest slab row
// among those items' classes, soft = shared-memory budget per CTA that keeps the intended number of CTAs on an SM.
template <int NT>
static cudaError_t launch_k2_part(const Plan& P, const WorkItem* d_items, uint32_t count, uint32_t max_stride, size_t soft, uint32_t nu, uint32_t nv,
                                  uint32_t NO, uint32_t NPT, int follows_sampler, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    // shared memory: 256 doubles of weights + the slabs (C and F of both sides) of as many points as fit; every CTA sizes its own
    // chunk from its class's row width, the launch only fixes the budget: `soft` unless the widest class cannot even stage one
    // quadrature row in it
    const size_t per_pt = (size_t)max_stride * 2 * sizeof(double);
    const size_t hard = (size_t)P.max_smem_optin - 1024;
    const uint32_t npts = nu * nv;
    const size_t fixed = 256 * sizeof(double);
    size_t smem = std::min(soft, fixed + (size_t)npts * per_pt);
    if ((smem - fixed) / per_pt < std::min<uint32_t>(npts, nv)) smem = std::min(hard, fixed + (size_t)std::min<uint32_t>(npts, nv) * per_pt);
    if ((smem - fixed) / per_pt == 0) return cudaErrorInvalidConfiguration;
    const uint32_t slab_doubles = (uint32_t)((smem - fixed) / sizeof(double));
    // Opt in to the device's full dynamic shared memory once per device (per NT), never less: the attribute is process-wide per
    // kernel, so raising it on demand from several host threads (multi-device calls) could lower it under another thread's launch.
    {
        static std::mutex mu;
        static bool done[64] = {};
        std::lock_guard<std::mutex> lk(mu);
        if (P.device >= 0 && P.device < 64 && !done[P.device]) {
            cudaError_t e = cudaFuncSetAttribute(k2_exact_kernel<K2_TILE_P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hard);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_exact_kernel<1, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hard);
            if (e != cudaSuccess) return e;
            done[P.device] = true;
        }
    }
    K2Args g{P.d_classes, P.d_lists, P.d_spec_i, P.d_spec_j, d_items, P.d_tabs, P.d_glq, P.d_V, NO, NPT, nu, nv, slab_doubles, follows_sampler, 0u, nullptr, 0};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(count); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = P.host.tile_p == 1 ? cudaLaunchKernelEx(&cfg, k2_exact_kernel<1, NT>, g) : cudaLaunchKernelEx(&cfg, k2_exact_kernel<K2_TILE_P, NT>, g);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// Shared memory of the warp-specialised integrator in front of its slab ring: weights, mbarriers, item contexts, the staging warp's
// table cache (4 tables) and the column table of the current pack.
static size_t ws_fixed_smem(const Plan& P, uint32_t NO, uint32_t NPT, uint32_t max_stride) {
    (void)P;
    return (256 + 2 * K2_WS_NBUF) * sizeof(double) + K2_WS_NCTX * sizeof(WsCtx) + 4 * (size_t)(3 * NO * ws_tab_stride(NPT)) * sizeof(double) +
           2 * (size_t)((max_stride + 3u) & ~3u) * sizeof(uint32_t);
}

// Persistent, warp-specialised launch over the packs [0, count) of an item list: two CTAs per SM, each with a ring of K2_WS_NBUF slab buffers.
static cudaError_t launch_k2_ws(const Plan& P, const WorkItem* d_items, const PackDesc* d_packs, uint32_t count, uint32_t max_stride, uint32_t nu, uint32_t nv,
                                uint32_t NO, uint32_t NPT, int follows_sampler, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    const size_t fixed = ws_fixed_smem(P, NO, NPT, max_stride);
    const size_t per_row = (size_t)max_stride * 2 * sizeof(double) * nv;     // chunks are whole quadrature rows
    const size_t hard = (size_t)P.max_smem_optin - 1024;
    const size_t smem = std::min<size_t>(hard, (size_t)K2_WS_SMEM_KB * 1024);      // two CTAs per SM
    if ((smem - fixed) / K2_WS_NBUF < per_row) return cudaErrorInvalidConfiguration;   // launch_k2_exact checks ws_fits() first
    const uint32_t buf_doubles = (uint32_t)(((smem - fixed) / K2_WS_NBUF / sizeof(double)) & ~(size_t)1);   // even: buffers stay 16-byte aligned
    {
        static std::mutex mu;
        static bool done[64] = {};
        std::lock_guard<std::mutex> lk(mu);
        if (P.device >= 0 && P.device < 64 && !done[P.device]) {
            cudaError_t e = cudaSuccess;
            auto optin = [&](auto kernel) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hard); };
            optin(k2_ws_kernel<K2_TILE_P, 1, 0>); optin(k2_ws_kernel<K2_TILE_P, 1, 1>); optin(k2_ws_kernel<K2_TILE_P, 1, 3>);
            optin(k2_ws_kernel<K2_TILE_P, 2, 0>); optin(k2_ws_kernel<K2_TILE_P, 2, 1>); optin(k2_ws_kernel<K2_TILE_P, 2, 3>);
            if (e != cudaSuccess) return e;
            done[P.device] = true;
        }
    }
    static const int by_subpartition = [] { const char* ev = std::getenv("FEM2D_K2_WS_ROLES"); return ev ? std::atoi(ev) : 1; }();   // tuning: 0 = staging warps by warp index
    K2Args g{P.d_classes, P.d_lists, P.d_spec_i, P.d_spec_j, d_items, P.d_tabs, P.d_glq, P.d_V, NO, NPT, nu, nv, buf_doubles, follows_sampler, (max_stride + 3u) & ~3u,
             P.d_class_geom, by_subpartition};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(std::min<uint32_t>(count, (uint32_t)(K2_MIN_CTAS * std::max(P.sm_count, 1)))); cfg.blockDim = dim3(K2_WS_THREADS);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // scales folded into the weights: the ratios only (1), ratios and max(det) (3), none (0)
    const uint32_t fold = (P.host.ws_fold & 3u) == 3u ? 3u : (P.host.ws_fold & 1u);
    cudaError_t e;
#define FEM2D_WS_LAUNCH(PRODW, FOLDV) e = cudaLaunchKernelEx(&cfg, k2_ws_kernel<K2_TILE_P, PRODW, FOLDV>, g, d_packs, count, P.d_work_counter)
    if (P.host.ws_prod == 2) { if (fold == 3u) FEM2D_WS_LAUNCH(2, 3); else if (fold == 1u) FEM2D_WS_LAUNCH(2, 1); else FEM2D_WS_LAUNCH(2, 0); }
    else { if (fold == 3u) FEM2D_WS_LAUNCH(1, 3); else if (fold == 1u) FEM2D_WS_LAUNCH(1, 1); else FEM2D_WS_LAUNCH(1, 0); }
#undef FEM2D_WS_LAUNCH
    return e != cudaSuccess ? e : cudaGetLastError();
}

// The warp-specialised integrator stages whole quadrature rows: one row (nv points) of the widest class must fit a ring buffer.
static bool ws_fits(const Plan& P, uint32_t max_stride, uint32_t nv, uint32_t NO, uint32_t NPT) {
    const size_t fixed = ws_fixed_smem(P, NO, NPT, max_stride);
    const size_t smem = std::min<size_t>((size_t)P.max_smem_optin - 1024, (size_t)K2_WS_SMEM_KB * 1024);
    return smem > fixed && (smem - fixed) / K2_WS_NBUF >= (size_t)max_stride * 2 * sizeof(double) * nv;
}

// Without packs (latency shape, FEM2D_K2_WS=0) the items are ordered by size (largest first); the first n_big of them run in K2_THREADS-wide CTAs, the rest in K2_SMALL_THREADS-wide
// ones.  The small-item grid goes first: its CTAs (8 per SM) fill the whole machine for about one wave, and the big-item CTAs move
// in as they drain (measured on cfg 4: 0.464 -> 0.452 ms against big-first, where the small grid ran in the big grid's tail at low
// occupancy).  Both launches carry the programmatic-dependent-launch attribute: the second grid starts once every CTA of the first has
// passed its wait for the sampler, runs alongside it and waits for its completion before exiting, so that whoever waits on the
// second grid (the scatter kernel) transitively waits on the first.
cudaError_t launch_k2_exact(const Plan& P, const WorkItem* d_items, uint32_t n_items, const PackDesc* d_packs, const ItemSplit& sp,
                            uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st, uint32_t* launches) {
    if (n_items == 0) return cudaSuccess;
    if (sp.n_packs) {
        // throughput shape: every item runs in the persistent warp-specialised grid, small items packed several to a round.  Fallback when
        // one quadrature row of the widest pack does not fit a ring buffer (very high orders, very many GLQ points): k2_exact_kernel
        // with 256-thread CTAs for every item (it chunks by points, not rows).
        if (ws_fits(P, sp.stride_pack, nv, NO, NPT)) {
            if (launches) (*launches)++;
            return launch_k2_ws(P, d_items, d_packs, sp.n_packs, sp.stride_pack, nu, nv, NO, NPT, 1, st);
        }
        if (launches) (*launches)++;
        return launch_k2_part<K2_THREADS>(P, d_items, n_items, std::max(sp.stride_big, sp.stride_small), (K2_MIN_CTAS == 2 ? 100 : 216 / K2_MIN_CTAS) * 1024, nu, nv, NO, NPT, 1, st);
    }
    const uint32_t n_big = std::min(sp.n_big, n_items);
    const uint32_t n_small = n_items - n_big;
    // small CTAs: 1/8 of an SM's shared memory each; big CTAs: prefer <= ~100 KB of shared memory so two share an SM
    cudaError_t e = launch_k2_part<K2_SMALL_THREADS>(P, d_items + n_big, n_small, sp.stride_small, 27 * 1024, nu, nv, NO, NPT, 1, st);
    if (e != cudaSuccess) return e;
    if (launches && n_small) (*launches)++;
    e = launch_k2_part<K2_THREADS>(P, d_items, n_big, sp.stride_big, (K2_MIN_CTAS == 2 ? 100 : 216 / K2_MIN_CTAS) * 1024, nu, nv, NO, NPT, n_small == 0, st);
    if (launches && n_big) (*launches)++;
    return e;
}

cudaError_t ws_profile(unsigned long long out[16], int reset) {
#ifdef FEM2D_WS_PROFILE
    cudaError_t e = cudaMemcpyFromSymbol(out, g_ws_prof, 16 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) { unsigned long long z[16] = {}; e = cudaMemcpyToSymbol(g_ws_prof, z, sizeof(z)); }
    return e;
#else
    for (int k = 0; k < 16; k++) out[k] = 0;
    (void)reset;
    return cudaSuccess;
#endif
}

cudaError_t fp64_peak(int kind, double* gflops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    double* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a);
        if (kind == 0) fp64_peak_kernel<0><<<blocks, threads>>>(d, iters); else fp64_peak_kernel<1><<<blocks, threads>>>(d, iters);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    if (e != cudaSuccess) return e;
    *gflops = (double)blocks * threads * iters * 8.0 * 2.0 / (best * 1e-3) / 1e9;   // 2 flops per (fma | mul+add)
    return cudaGetLastError();
}

}  // namespace fem2d
