// Re-ordered ("fast") integrators: mathematically equal to the reference integrals, NOT bit-identical (different summation
// order, FMA contraction allowed).  They are judged with the scale-aware tolerance (1e-12 relative, floor 1e-14 * max|M|),
// see DESIGN.md section 2.  The bit-faithful kernel (kernels_exact.cu) remains the referee and the product default.
//
// FEM2D_MODE_SUMFACT -- sum factorisation.  The Jacobian is constant per Elem (element.rs:33-50), so every
// integrand separates:  sum_m sum_n w_m w_n X_i(m) Y_i'(m) Z_j(n) W_j'(n)  =  Gu[X,Y][i,i'] * Gv[Z,W][j,j'],
// with 1-D Gram matrices  Gu[X,Y][i,i'] = sum_m w_m X^P_i(m) Y^Q_i'(m)  (P tables at the RBS-mapped points, Q tables unscaled).
//   A(U,U) =  (1/mu) jiuP jiuQ su ratio_uv * Gu[N ,N ] * Gv[T',T']        B(U,U) = eps (su sv) jiuP jiuQ maxdet * Gu[N,N] * Gv[T,T]
//   A(V,V) =  (1/mu) jivP jivQ sv ratio_vu * Gu[T',T'] * Gv[N ,N ]        B(V,V) = eps (su sv) jivP jivQ maxdet * Gu[T,T] * Gv[N,N]
//   A(U,V) = -(1/mu) jiuP jivQ sv          * Gu[N ,T'] * Gv[T',N ]        B(U,V) = B(V,U) = 0
//   A(V,U) = -(1/mu) jivP jiuQ su          * Gu[T',N ] * Gv[N ,T']
// (P's derivative is scaled by Q's para_scale = (1,1), Q's by P's = (su,sv): integrals.rs:44,47; basis.rs:235-252.)
// The integrator then is a pure streaming kernel: 2 loads + 3 multiplies per value, 16 B written per pair -> HBM bound.
//
// FEM2D_MODE_DMMA -- Phi^T W Phi contraction of the sampled curl / value slabs on FP64 tensor-core tiles
// (mma.sync.aligned.m8n8k4.row.col.f64), the dense form named by BASELINE.json north_star (2).  It does O(n^2 Q) work where
// sum factorisation does O(n^2): it is the right tool only for non-separable integrands (curvilinear Elements, which the
// reference does not implement: element.rs:31 TODO); it is kept as the measured alternative.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "device_plan.hpp"

namespace fem2d {
namespace {

enum { G_NN = 0, G_TDTD = 1, G_NTD = 2, G_TDN = 3, G_TT = 4, G_KINDS = 5 };

// One CTA per Gram set: G[kind][i][i'] over the axis' points.  Table layout as written by K1: [arr][order][point], arr 0 N, 1 N', 2 T, 3 T'.
__global__ void gram_kernel(const GramDesc* __restrict__ grams, const double* __restrict__ tabs, const double* __restrict__ glq,
                            double* __restrict__ out, uint32_t NO, uint32_t NPT, uint32_t nu, uint32_t nv) {
    const GramDesc g = grams[blockIdx.x];
    const uint32_t np = g.axis ? nv : nu;
    const double* w = glq + (g.axis ? 384 : 128);
    const double* tp = tabs + (size_t)g.tabP * 4 * NO * NPT;
    const double* tq = tabs + (size_t)g.tabQ * 4 * NO * NPT;
    double* o = out + (size_t)blockIdx.x * G_KINDS * NO * NO;
    // array index of the P factor / Q factor per kind: N=0, T=2, T'=3
    const int ap[G_KINDS] = {0, 3, 0, 3, 2}, aq[G_KINDS] = {0, 3, 3, 0, 2};
    for (uint32_t t = threadIdx.x; t < G_KINDS * NO * NO; t += blockDim.x) {
        const uint32_t kind = t / (NO * NO), r = t - kind * NO * NO, i = r / NO, k = r - i * NO;
        const double* x = tp + ((size_t)ap[kind] * NO + i) * NPT;
        const double* y = tq + ((size_t)aq[kind] * NO + k) * NPT;
        double s = 0.0;
        for (uint32_t p = 0; p < np; p++) s += w[p] * x[p] * y[p];
        o[t] = s;
    }
}

// Per-class constants shared by both re-ordered integrators.
struct ClassConsts {
    double kA_uu, kA_vv, kA_uv, kA_vu, kB_uu, kB_vv;
};
__device__ __forceinline__ ClassConsts class_consts(const ClassDesc& c) {
    const double detP = c.dxP * c.dyP, detQ = c.dxQ * c.dyQ;
    const double jiuP = c.dyP / detP, jivP = c.dxP / detP, jiuQ = c.dyQ / detQ, jivQ = c.dxQ / detQ;
    const bool pge = detP >= detQ;
    const double ratio_uv = pge ? c.dxP / c.dyP : c.dxQ / c.dyQ, ratio_vu = pge ? c.dyP / c.dxP : c.dyQ / c.dxQ;
    const double maxdet = detP > detQ ? detP : detQ;
    const double coefA = 1.0 / c.mu, coefB = c.eps * (c.su * c.sv);
    ClassConsts k;
    k.kA_uu = coefA * jiuP * jiuQ * c.su * ratio_uv; k.kA_vv = coefA * jivP * jivQ * c.sv * ratio_vu;
    k.kA_uv = -coefA * jiuP * jivQ * c.sv;           k.kA_vu = -coefA * jivP * jiuQ * c.su;
    k.kB_uu = coefB * jiuP * jiuQ * maxdet;          k.kB_vv = coefB * jivP * jivQ * maxdet;
    return k;
}

struct SFArgs {
    const ClassDesc* classes; const ListDesc* lists; const uint8_t* spec_i; const uint8_t* spec_j; const double* gram; double2* V;
    uint32_t NO;
};

// One CTA per class; threads stream over the nP x nQ pairs (consecutive threads -> consecutive q -> coalesced 16-byte stores).
__global__ void __launch_bounds__(256) k2_sumfact_kernel(const SFArgs g) {
    const ClassDesc c = g.classes[blockIdx.x];
    const ListDesc LP = g.lists[c.listP], LQ = g.lists[c.listQ];
    const uint32_t nP = LP.n, nUP = LP.nU, nQ = LQ.n, nUQ = LQ.nU, NO = g.NO, GS = NO * NO;
    const ClassConsts kc = class_consts(c);
    const double* Gu = g.gram + (size_t)c.gramU * G_KINDS * GS;
    const double* Gv = g.gram + (size_t)c.gramV * G_KINDS * GS;
    double2* out = g.V + c.v_off;
    const uint32_t total = nP * nQ;
    for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
        const uint32_t a = t / nQ, b = t - a * nQ;
        const uint32_t i = g.spec_i[LP.off + a], j = g.spec_j[LP.off + a], k = g.spec_i[LQ.off + b], l = g.spec_j[LQ.off + b];
        const bool pu = a < nUP, qu = b < nUQ;
        const uint32_t uu = i * NO + k, vv = j * NO + l;
        double A, B = 0.0;
        if (pu && qu) { A = kc.kA_uu * Gu[G_NN * GS + uu] * Gv[G_TDTD * GS + vv]; B = kc.kB_uu * Gu[G_NN * GS + uu] * Gv[G_TT * GS + vv]; }
        else if (!pu && !qu) { A = kc.kA_vv * Gu[G_TDTD * GS + uu] * Gv[G_NN * GS + vv]; B = kc.kB_vv * Gu[G_TT * GS + uu] * Gv[G_NN * GS + vv]; }
        else if (pu) A = kc.kA_uv * Gu[G_NTD * GS + uu] * Gv[G_TDN * GS + vv];
        else A = kc.kA_vu * Gu[G_TDN * GS + uu] * Gv[G_NTD * GS + vv];
        out[t] = make_double2(A, B);
    }
}

cudaError_t ensure_gram(Plan& P, uint32_t NO, cudaStream_t st) {
    const size_t need = std::max<size_t>(P.host.grams.size(), 1) * G_KINDS * NO * NO;
    if (need > P.gram_capacity) {
        // an earlier call on this stream may still read the old buffer, and the block cache can hand it straight to another allocation
        const cudaError_t es = cudaStreamSynchronize(st);
        if (es != cudaSuccess) return es;
        fem2d::dev_free(P.d_gram, st); P.d_gram = nullptr; P.gram_capacity = 0;
        cudaError_t e = fem2d::dev_malloc((void**)&P.d_gram, need * sizeof(double), st);
        if (e != cudaSuccess) return e;
        P.gram_capacity = need;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------------------------ DMMA
// Block (class) matrices as dense contractions over the quadrature points:
//   A[a][b] = kA * sum_pt (sqrt(w_pt) curlP[a][pt]) * (sqrt(w_pt) curlQ[b][pt]),   B likewise with the value factors.
// curl / val are the per-function factors without the per-sample constants (folded into kA / kB per direction pair); the
// quadrature weight is split as sqrt(w) * sqrt(w) so that a local class (P == Q sample) stages ONE slab pair for both operands.
// Slabs are function-major [f][pt] with a point stride == 4 (mod 16), which makes the 8x4 / 4x8 fragment loads bank-conflict
// free.  Functions are grouped U first then V, each group padded to a multiple of 8 -> every 8x8 MMA tile is direction-homogeneous.
constexpr int DM_WARPS = 16;
constexpr int DM_TILES_PER_WARP = 6;
constexpr int DM_ITEM_TILES = DM_WARPS * DM_TILES_PER_WARP;

struct DmmaItem { uint32_t cls, tile_begin, tile_count, pad; };
struct DMArgs {
    const ClassDesc* classes; const ListDesc* lists; const uint8_t* spec_i; const uint8_t* spec_j; const DmmaItem* items;
    const double* tabs; const double* glq; double2* V;
    uint32_t NO, NPT, nu, nv, chunk_pts, PS;
};

__device__ __forceinline__ uint32_t pad8(uint32_t x) { return (x + 7u) & ~7u; }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(DM_WARPS * 32) k2_dmma_kernel(const DMArgs g) {
    extern __shared__ __align__(16) double smem[];
    const DmmaItem it = g.items[blockIdx.x];
    const ClassDesc c = g.classes[it.cls];
    const ListDesc LP = g.lists[c.listP], LQ = g.lists[c.listQ];
    const uint32_t nP = LP.n, nUP = LP.nU, nQ = LQ.n, nUQ = LQ.nU;
    const uint32_t rowsP = pad8(nUP) + pad8(nP - nUP), rowsQ = c.local ? rowsP : pad8(nUQ) + pad8(nQ - nUQ);
    const uint32_t nu = g.nu, nv = g.nv, npts = nu * nv, PS = g.PS, chunk = g.chunk_pts;
    double* s_sw = smem;                                   // [128 + 128] sqrt(u_w), sqrt(v_w)
    double* s_CP = smem + 256;                             // [rowsP][PS]
    double* s_FP = s_CP + (size_t)rowsP * PS;
    double* s_CQ = c.local ? s_CP : s_FP + (size_t)rowsP * PS;
    double* s_FQ = c.local ? s_FP : s_CQ + (size_t)rowsQ * PS;
    for (uint32_t k = threadIdx.x; k < nu; k += blockDim.x) s_sw[k] = sqrt(g.glq[128 + k]);
    for (uint32_t k = threadIdx.x; k < nv; k += blockDim.x) s_sw[128 + k] = sqrt(g.glq[384 + k]);

    const double* tPu = g.tabs + (size_t)c.tabPu * 4 * g.NO * g.NPT;
    const double* tPv = g.tabs + (size_t)c.tabPv * 4 * g.NO * g.NPT;
    const double* tQu = g.tabs + (size_t)c.tabQu * 4 * g.NO * g.NPT;
    const double* tQv = g.tabs + (size_t)c.tabQv * 4 * g.NO * g.NPT;
    const uint32_t AS = g.NO * g.NPT;

    // tile grid: row tiles = U tiles then V tiles of P, col tiles likewise for Q
    const uint32_t nRtU = pad8(nUP) / 8, nCtU = pad8(nUQ) / 8, nCt = rowsQ / 8;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tile_rc[DM_TILES_PER_WARP];   // row tile << 16 | col tile
    double accA[DM_TILES_PER_WARP][2], accB[DM_TILES_PER_WARP][2];
    uint32_t n_my = 0;
#pragma unroll
    for (uint32_t k = 0; k < DM_TILES_PER_WARP; k++) {
        accA[k][0] = accA[k][1] = accB[k][0] = accB[k][1] = 0.0;
        uint32_t t = warp + k * DM_WARPS;
        tile_rc[k] = 0;
        if (t < it.tile_count) {
            t += it.tile_begin;
            uint32_t rt, ct;
            if (c.local) { rt = 0; while (t >= nCt - rt) { t -= nCt - rt; rt++; } ct = rt + t; }
            else { rt = t / nCt; ct = t - rt * nCt; }
            tile_rc[k] = rt << 16 | ct; n_my = k + 1;
        }
    }

    for (uint32_t pt0 = 0; pt0 < npts; pt0 += chunk) {
        const uint32_t cn = min(chunk, npts - pt0), cn4 = (cn + 3u) & ~3u;
        __syncthreads();
        // ---- stage: task = (slab row, quadrature row m); the n loop runs inside.  Padding rows / points are zero.
        const uint32_t m_lo = pt0 / nv, m_hi = (pt0 + cn - 1) / nv;
        for (int side = 0; side < (c.local ? 1 : 2); side++) {
            const uint32_t rows = side ? rowsQ : rowsP, nF = side ? nQ : nP, nUF = side ? nUQ : nUP, loff = side ? LQ.off : LP.off;
            const double* tu = side ? tQu : tPu; const double* tv = side ? tQv : tPv;
            double* sC = side ? s_CQ : s_CP; double* sF = side ? s_FQ : s_FP;
            const uint32_t padU = pad8(nUF);
            const uint32_t ntask = rows * (m_hi - m_lo + 1);
            for (uint32_t t = threadIdx.x; t < ntask; t += blockDim.x) {
                const uint32_t mi = t / rows, r = t - mi * rows, m = m_lo + mi;
                const uint32_t n_lo = (m == m_lo) ? pt0 - m_lo * nv : 0u;
                const uint32_t n_hi = (m == m_hi) ? pt0 + cn - 1 - m_hi * nv : nv - 1;
                double* dC = sC + (size_t)r * PS + (m * nv + n_lo - pt0);
                double* dF = sF + (size_t)r * PS + (m * nv + n_lo - pt0);
                const double swm = s_sw[m];
                if (r < nUF) {
                    const uint32_t i = g.spec_i[loff + r], j = g.spec_j[loff + r];
                    const double Ni = swm * tu[(0 * g.NO + i) * g.NPT + m];
                    const double* Tj = tv + 2 * AS + j * g.NPT; const double* Tdj = tv + 3 * AS + j * g.NPT;
                    for (uint32_t n = n_lo; n <= n_hi; n++) { const double w = s_sw[128 + n] * Ni; *dC++ = w * Tdj[n]; *dF++ = w * Tj[n]; }   // N_i T'_j | N_i T_j
                } else if (r >= padU && r - padU < nF - nUF) {
                    const uint32_t a = nUF + (r - padU);
                    const uint32_t i = g.spec_i[loff + a], j = g.spec_j[loff + a];
                    const double Ti = swm * tu[2 * AS + i * g.NPT + m], Tdi = swm * tu[3 * AS + i * g.NPT + m];
                    const double* Nj = tv + (0 * g.NO + j) * g.NPT;
                    for (uint32_t n = n_lo; n <= n_hi; n++) { const double w = s_sw[128 + n] * Nj[n]; *dC++ = w * Tdi; *dF++ = w * Ti; }     // T'_i N_j | T_i N_j
                } else {
                    for (uint32_t n = n_lo; n <= n_hi; n++) { *dC++ = 0.0; *dF++ = 0.0; }
                }
                if (m == m_hi) for (uint32_t p = cn; p < cn4; p++) { sC[(size_t)r * PS + p] = 0.0; sF[(size_t)r * PS + p] = 0.0; }
            }
        }
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < DM_TILES_PER_WARP; k++) {
            if (k >= n_my) break;
            const uint32_t rt = tile_rc[k] >> 16, ct = tile_rc[k] & 0xffffu;
            const bool same = (rt < nRtU) == (ct < nCtU);
            const size_t ao = (size_t)(rt * 8 + (lane >> 2)) * PS + (lane & 3), bo = (size_t)(ct * 8 + (lane >> 2)) * PS + (lane & 3);
            const double* ap = s_CP + ao; const double* bp = s_CQ + bo;
            double a0 = accA[k][0], a1 = accA[k][1];
            if (same) {
                const double* afp = s_FP + ao; const double* bfp = s_FQ + bo;
                double b0 = accB[k][0], b1 = accB[k][1];
#pragma unroll 4
                for (uint32_t p = 0; p < cn4; p += 4) { dmma884(a0, a1, ap[p], bp[p]); dmma884(b0, b1, afp[p], bfp[p]); }
                accB[k][0] = b0; accB[k][1] = b1;
            } else {
#pragma unroll 4
                for (uint32_t p = 0; p < cn4; p += 4) dmma884(a0, a1, ap[p], bp[p]);
            }
            accA[k][0] = a0; accA[k][1] = a1;
        }
    }
    // ---- epilogue: per-direction constants, write the 8x8 tile (lane holds C[lane/4][2*(lane%4) + {0,1}])
    const ClassConsts kc = class_consts(c);
    double2* out = g.V + c.v_off;
#pragma unroll
    for (uint32_t k = 0; k < DM_TILES_PER_WARP; k++) {
        if (k >= n_my) break;
        const uint32_t rt = tile_rc[k] >> 16, ct = tile_rc[k] & 0xffffu;
        const bool rowU = rt < nRtU, colU = ct < nCtU;
        const uint32_t a = rowU ? rt * 8 + (lane >> 2) : nUP + (rt - nRtU) * 8 + (lane >> 2);
        const uint32_t b = colU ? ct * 8 + 2 * (lane & 3) : nUQ + (ct - nCtU) * 8 + 2 * (lane & 3);
        const uint32_t a_end = rowU ? nUP : nP, b_end = colU ? nUQ : nQ;
        const double kA = rowU ? (colU ? kc.kA_uu : kc.kA_uv) : (colU ? kc.kA_vu : kc.kA_vv);
        const double kB = rowU == colU ? (rowU ? kc.kB_uu : kc.kB_vv) : 0.0;
        if (a < a_end) {
            if (b < b_end) out[(size_t)a * nQ + b] = make_double2(kA * accA[k][0], kB * accB[k][0]);
            if (b + 1 < b_end) out[(size_t)a * nQ + b + 1] = make_double2(kA * accA[k][1], kB * accB[k][1]);
        }
    }
}

}  // namespace

cudaError_t launch_k2_sumfact(Plan& P, uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st, uint32_t* launches) {
    if (P.host.classes.empty()) return cudaSuccess;
    cudaError_t e = ensure_gram(P, NO, st);
    if (e != cudaSuccess) return e;
    gram_kernel<<<(unsigned)P.host.grams.size(), 128, 0, st>>>(P.d_grams, P.d_tabs, P.d_glq, P.d_gram, NO, NPT, nu, nv);
    if (launches) (*launches)++;
    SFArgs g{P.d_classes, P.d_lists, P.d_spec_i, P.d_spec_j, P.d_gram, P.d_V, NO};
    k2_sumfact_kernel<<<(unsigned)P.host.classes.size(), 256, 0, st>>>(g);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_k2_dmma(Plan& P, uint32_t nu, uint32_t nv, uint32_t NO, uint32_t NPT, cudaStream_t st, uint32_t* launches) {
    if (P.host.classes.empty()) return cudaSuccess;
    auto p8 = [](uint32_t x) { return (x + 7u) & ~7u; };
    // tile work items (built once per plan)
    if (!P.d_dmma_items) {
        std::vector<DmmaItem> items;
        for (uint32_t ci = 0; ci < P.host.classes.size(); ci++) {
            const ClassDesc& c = P.host.classes[ci];
            const ListDesc& LP = P.host.lists[c.listP]; const ListDesc& LQ = P.host.lists[c.listQ];
            const uint32_t nRt = (p8(LP.nU) + p8(LP.n - LP.nU)) / 8, nCt = (p8(LQ.nU) + p8(LQ.n - LQ.nU)) / 8;
            const uint32_t tiles = c.local ? nRt * (nRt + 1) / 2 : nRt * nCt;
            const uint32_t n_items = (tiles + DM_ITEM_TILES - 1) / DM_ITEM_TILES;
            for (uint32_t k = 0; k < n_items; k++) {
                const uint32_t b = (uint32_t)((uint64_t)tiles * k / n_items), en = (uint32_t)((uint64_t)tiles * (k + 1) / n_items);
                if (en > b) items.push_back(DmmaItem{ci, b, en - b, 0});
            }
        }
        P.n_dmma_items = (uint32_t)items.size();
        cudaError_t e = fem2d::dev_malloc((void**)&P.d_dmma_items, std::max<size_t>(items.size(), 1) * sizeof(DmmaItem), st);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(P.d_dmma_items, items.data(), items.size() * sizeof(DmmaItem), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);   // `items` is a local vector
        if (e != cudaSuccess) return e;
    }
    uint32_t max_rows = 0;   // slab rows of the largest class (local classes share one slab pair between P and Q)
    for (const ClassDesc& c : P.host.classes) {
        const ListDesc& LP = P.host.lists[c.listP]; const ListDesc& LQ = P.host.lists[c.listQ];
        uint32_t r = p8(LP.nU) + p8(LP.n - LP.nU);
        if (!c.local) r += p8(LQ.nU) + p8(LQ.n - LQ.nU);
        max_rows = std::max(max_rows, r);
    }
    const uint32_t npts = nu * nv;
    const size_t fixed = 256 * sizeof(double);
    const size_t soft = 100 * 1024, hard = (size_t)P.max_smem_optin - 2048;
    // point stride == 4 (mod 16); chunk = as many points (multiple of 4) as fit, preferring <= ~100 KB (two CTAs per SM)
    auto ps_of = [](uint32_t ch) { return ((ch + 15u) & ~15u) + 4u; };
    auto bytes_of = [&](uint32_t ch) { return fixed + (size_t)max_rows * ps_of(ch) * 2 * sizeof(double); };
    uint32_t chunk = (npts + 3u) & ~3u;
    if (bytes_of(chunk) > soft) {
        uint32_t c2 = chunk;
        while (c2 > 16 && bytes_of(c2) > soft) c2 -= 4;
        if (bytes_of(c2) <= soft && c2 * 2 >= chunk) chunk = c2;      // at most two chunks within the soft budget
        while (chunk > 4 && bytes_of(chunk) > hard) chunk -= 4;
    }
    if (bytes_of(chunk) > hard) return cudaErrorInvalidConfiguration;
    const uint32_t PS = ps_of(chunk);
    const size_t smem = bytes_of(chunk);
    cudaError_t e = cudaFuncSetAttribute(k2_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hard);   // always the full opt-in: never lowered under another thread's launch
    if (e != cudaSuccess) return e;
    DMArgs g{P.d_classes, P.d_lists, P.d_spec_i, P.d_spec_j, (const DmmaItem*)P.d_dmma_items, P.d_tabs, P.d_glq, P.d_V, NO, NPT, nu, nv, chunk, PS};
    k2_dmma_kernel<<<P.n_dmma_items, DM_WARPS * 32, smem, st>>>(g);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace fem2d
