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

#include "basis_device.cuh"
#include "device_plan.hpp"

namespace fem2d {

namespace {

// ---------------------------------------------------------------------------------------------------------------- K1
// One CTA per table, one thread per point.  Output layout: out[((arr * NO) + order) * NPT + point], arr: 0 N, 1 N', 2 T, 3 T'.
__global__ void k1_tables_kernel(const TableDesc* __restrict__ tabs, double* __restrict__ out, uint32_t NO, uint32_t NPT,
                                 const double* __restrict__ glq, uint32_t nu, uint32_t nv, uint32_t i_max, uint32_t j_max, int basis) {
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
};

__device__ __forceinline__ uint32_t pad4(uint32_t x) { return (x + 3u) & ~3u; }

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
    k1_tables_kernel<<<(unsigned)P.host.tables.size(), 64, 0, st>>>(P.d_tables, P.d_tabs, NO, NPT, P.d_glq, nu, nv, P.host.i_max, P.host.j_max, basis_kind);
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
    static thread_local size_t smem_set[64] = {};   // per device: largest dynamic shared memory size already opted in (per NT)
    if (P.device >= 0 && P.device < 64 && smem > smem_set[P.device]) {
        cudaError_t e = cudaFuncSetAttribute(k2_exact_kernel<K2_TILE_P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_exact_kernel<1, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set[P.device] = smem;
    }
    K2Args g{P.d_classes, P.d_lists, P.d_spec_i, P.d_spec_j, d_items, P.d_tabs, P.d_glq, P.d_V, NO, NPT, nu, nv, slab_doubles, follows_sampler};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(count); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = P.host.tile_p == 1 ? cudaLaunchKernelEx(&cfg, k2_exact_kernel<1, NT>, g) : cudaLaunchKernelEx(&cfg, k2_exact_kernel<K2_TILE_P, NT>, g);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// Items are ordered by size (largest first); the first n_big of them run in K2_THREADS-wide CTAs, the rest in K2_SMALL_THREADS-wide
// ones.  The small-item grid goes first: its CTAs (8 per SM) fill the whole machine for about one wave, and the big-item CTAs move
// in as they drain (measured on cfg 4: 0.464 -> 0.452 ms against big-first, where the small grid ran in the big grid's tail at low
// occupancy).  Both launches carry the programmatic-dependent-launch attribute: the second grid starts once every CTA of the first has
// passed its wait for the sampler, runs alongside it and waits for its completion before exiting, so that whoever waits on the
// second grid (the scatter kernel) transitively waits on the first.
cudaError_t launch_k2_exact(const Plan& P, const WorkItem* d_items, uint32_t n_items, uint32_t n_big, uint32_t stride_big, uint32_t stride_small,
                            uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st, uint32_t* launches) {
    if (n_items == 0) return cudaSuccess;
    n_big = std::min(n_big, n_items);
    const uint32_t n_small = n_items - n_big;
    // small CTAs: 1/8 of an SM's shared memory each; big CTAs: prefer <= ~100 KB of shared memory so two share an SM
    cudaError_t e = launch_k2_part<K2_SMALL_THREADS>(P, d_items + n_big, n_small, stride_small, 27 * 1024, nu, nv, NO, NPT, 1, st);
    if (e != cudaSuccess) return e;
    if (launches && n_small) (*launches)++;
    e = launch_k2_part<K2_THREADS>(P, d_items, n_big, stride_big, (K2_MIN_CTAS == 2 ? 100 : 216 / K2_MIN_CTAS) * 1024, nu, nv, NO, NPT, n_small == 0, st);
    if (launches && n_big) (*launches)++;
    return e;
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
