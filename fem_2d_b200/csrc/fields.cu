// UniformFieldSpace::xy_fields (fields.rs:63-127) on the GPU -- the caller-side "next" row of SURVEY.md section 8f.
//
// For every leaf ("shell") Elem and every ancestor (self first, then towards the base, mesh.rs:520-541) the ancestor's basis
// functions are sampled over the leaf on a uniform d x d grid (HierCurlBasisFn::defined_over(anc, Some(leaf), uniform points),
// fields.rs:93-99) and accumulated as  value = f_u|f_v(i, j, [m, n]) * solution[dof]  (fields.rs:101-115), in the reference's
// order (ancestors outer, the Elem's BasisSpec list in its stored order inner) with separately rounded operations, so the
// result is bit-identical to the reference algorithm.  One CTA per leaf, one thread per grid point.
// COMPILED WITH -fmad=false.  The x (y) component only receives U (V) directed functions: the other component of
// jac_inv.u / jac_inv.v is -0.0 (space.rs:142-147) and contributes signed zeros only.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/fem2d.h"
#include "basis_device.cuh"
#include "device_plan.hpp"

namespace fem2d {
namespace {

struct FieldArgs {
    const uint32_t* leaf_ids; const int32_t* elem_parent; const uint8_t* elem_loc; const double* elem_dx; const double* elem_dy;
    const uint32_t* bs_off; const uint8_t* bs_i; const uint8_t* bs_j; const uint8_t* bs_dir; const uint32_t* bs_dof;
    const double* solution; double* x_out; double* y_out;
    uint32_t d, i_max, j_max; int basis;
};

// child sub-range (h_refinement.rs:247-279)
__device__ __forceinline__ void dev_sub_range(uint8_t loc, double* r) {
    const double mu = (r[0] + r[1]) / 2.0, mv = (r[2] + r[3]) / 2.0;
    const bool west = (loc == FEM2D_LOC_SW || loc == FEM2D_LOC_NW || loc == FEM2D_LOC_W), east = (loc == FEM2D_LOC_SE || loc == FEM2D_LOC_NE || loc == FEM2D_LOC_E);
    const bool south = (loc == FEM2D_LOC_SW || loc == FEM2D_LOC_SE || loc == FEM2D_LOC_S), north = (loc == FEM2D_LOC_NW || loc == FEM2D_LOC_NE || loc == FEM2D_LOC_N);
    if (east) r[0] = mu;
    if (west) r[1] = mu;
    if (north) r[2] = mv;
    if (south) r[3] = mv;
}

__global__ void __launch_bounds__(256) xy_fields_kernel(const FieldArgs g) {
    const uint32_t leaf = g.leaf_ids[blockIdx.x];
    const uint32_t d = g.d, npt = d * d;
    // uniform_range(-1, 1, d) (fields.rs:407-410): step = (max - min) / (n - 1); p_i = i * step + min
    const double step = (1.0 - (-1.0)) / (double)(d - 1);
    for (uint32_t pt = threadIdx.x; pt < npt; pt += blockDim.x) {
        const uint32_t m = pt / d, n = pt - m * d;
        const double pu = (double)m * step + (-1.0), pv = (double)n * step + (-1.0);
        double xs = 0.0, ys = 0.0;
        uint32_t depth = 0;          // number of HRefLocs between the current ancestor and the leaf
        for (int32_t anc = (int32_t)leaf; anc >= 0; anc = g.elem_parent[anc]) {
            // relative_parametric_range (elem.rs:170-188): fold from the child of `anc` down to the leaf
            double su = 1.0, ou = 0.0, sv = 1.0, ov = 0.0;
            if (depth > 0) {
                double r[4] = {-1.0, 1.0, -1.0, 1.0};
                uint8_t locs[64];
                uint32_t k = 0;
                for (int32_t e = (int32_t)leaf; e != anc; e = g.elem_parent[e]) locs[k++] = g.elem_loc[e];
                for (int32_t q = (int32_t)k - 1; q >= 0; q--) dev_sub_range(locs[q], r);
                su = (r[1] - r[0]) / 2.0; ou = (r[1] + r[0]) / 2.0;   // scale_gauss_quad_points (glq.rs:238-249)
                sv = (r[3] - r[2]) / 2.0; ov = (r[3] + r[2]) / 2.0;
            }
            const uint32_t b0 = g.bs_off[anc], b1 = g.bs_off[anc + 1];
            if (b1 > b0) {
                const double xu = depth > 0 ? pu * su + ou : pu, xv = depth > 0 ? pv * sv + ov : pv;
                double Nu[21], Tu[21], Nv[21], Tv[21];
                basis_at_point(g.basis, g.i_max, xu, [&](int arr, uint32_t q, double val) { if (arr == 0) Nu[q] = val; else if (arr == 2) Tu[q] = val; });
                basis_at_point(g.basis, g.j_max, xv, [&](int arr, uint32_t q, double val) { if (arr == 0) Nv[q] = val; else if (arr == 2) Tv[q] = val; });
                // constant Jacobian of the ANCESTOR's own extent (basis.rs:400; element.rs:46-49), inverse as in space.rs:142-147
                const double dx = g.elem_dx[anc], dy = g.elem_dy[anc];
                const double det = dx * dy - 0.0 * 0.0;
                const double jiu = dy / det, jiv = dx / det;
                for (uint32_t s = b0; s < b1; s++) {
                    const uint32_t i = g.bs_i[s], j = g.bs_j[s];
                    const double sol = g.solution[g.bs_dof[s]];
                    if (g.bs_dir[s] == 0) xs = xs + ((jiu * Nu[i]) * Tv[j]) * sol;      // f_u (basis.rs:225-227) * solution
                    else ys = ys + ((jiv * Tu[i]) * Nv[j]) * sol;                       // f_v (basis.rs:230-232) * solution
                }
            }
            depth++;
        }
        g.x_out[(size_t)blockIdx.x * npt + pt] = xs;   // x_values[m][n] (fields.rs:111)
        g.y_out[(size_t)blockIdx.x * npt + pt] = ys;
    }
}

template <class T>
cudaError_t up(T** dst, const T* src, size_t n) {
    cudaError_t e = dev_malloc((void**)dst, n * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, nullptr);
}

}  // namespace
}  // namespace fem2d

extern "C" int fem2d_xy_fields(const fem2d_domain_view* v, int device, int basis_kind, uint32_t density, const double* solution,
                               uint64_t leaf_capacity, uint64_t* n_leaves, uint32_t* leaf_ids, double* x_out, double* y_out) {
    using namespace fem2d;
    if (!v || !solution || !n_leaves) return FEM2D_ERR_BAD_ARGUMENT;
    if (density < 2 || density > 64) return FEM2D_ERR_UNSUPPORTED;
    if (basis_kind != FEM2D_BASIS_HIER_POLY && basis_kind != FEM2D_BASIS_HIER_MAX_ORTHO) return FEM2D_ERR_UNSUPPORTED;
    if (v->i_max > 20 || v->j_max > 20 || (basis_kind == FEM2D_BASIS_HIER_MAX_ORTHO && (v->i_max > 12 || v->j_max > 12))) return FEM2D_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return FEM2D_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) return FEM2D_ERR_BAD_ARGUMENT;
    const uint32_t ne = v->n_elems;
    // leaves = Elems without children, in id order (fields.rs:82)
    std::vector<uint8_t> has_child(ne, 0);
    for (uint32_t e = 0; e < ne; e++) if (v->elem_parent[e] >= 0) has_child[v->elem_parent[e]] = 1;
    std::vector<uint32_t> leaves;
    for (uint32_t e = 0; e < ne; e++) if (!has_child[e]) leaves.push_back(e);
    *n_leaves = leaves.size();
    if (leaves.size() > leaf_capacity || !leaf_ids || !x_out || !y_out) return leaves.size() > leaf_capacity ? FEM2D_ERR_BAD_ARGUMENT : FEM2D_ERR_BAD_ARGUMENT;
    // per-Elem Jacobian through the planner's geometry (same arithmetic as the assembly path)
    std::string err;
    std::vector<double> dx, dy;
    if (int st = elem_geometry(v, dx, dy, err)) return st;
    if (cudaSetDevice(device) != cudaSuccess) return FEM2D_ERR_CUDA;
    dev_pool_init(device);
    const uint32_t nbs = v->bs_off[ne];
    const size_t npt = (size_t)density * density;
    uint32_t *d_leaf = nullptr, *d_off = nullptr, *d_dof = nullptr; int32_t* d_par = nullptr; uint8_t *d_loc = nullptr, *d_i = nullptr, *d_j = nullptr, *d_dir = nullptr;
    double *d_dx = nullptr, *d_dy = nullptr, *d_sol = nullptr, *d_x = nullptr, *d_y = nullptr;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    ok(up(&d_leaf, leaves.data(), leaves.size())); ok(up(&d_par, v->elem_parent, ne)); ok(up(&d_loc, v->elem_loc, ne));
    ok(up(&d_dx, dx.data(), ne)); ok(up(&d_dy, dy.data(), ne)); ok(up(&d_off, v->bs_off, ne + 1));
    ok(up(&d_i, v->bs_i, nbs ? nbs : 1)); ok(up(&d_j, v->bs_j, nbs ? nbs : 1)); ok(up(&d_dir, v->bs_dir, nbs ? nbs : 1)); ok(up(&d_dof, v->bs_dof, nbs ? nbs : 1));
    ok(up(&d_sol, solution, v->n_dofs ? v->n_dofs : 1));
    ok(dev_malloc((void**)&d_x, leaves.size() * npt * sizeof(double))); ok(dev_malloc((void**)&d_y, leaves.size() * npt * sizeof(double)));
    if (e == cudaSuccess && !leaves.empty()) {
        FieldArgs g{d_leaf, d_par, d_loc, d_dx, d_dy, d_off, d_i, d_j, d_dir, d_dof, d_sol, d_x, d_y, density, v->i_max, v->j_max, basis_kind};
        xy_fields_kernel<<<(unsigned)leaves.size(), 256>>>(g);
        ok(cudaGetLastError());
        ok(cudaMemcpyAsync(x_out, d_x, leaves.size() * npt * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
        ok(cudaMemcpyAsync(y_out, d_y, leaves.size() * npt * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
    }
    ok(cudaStreamSynchronize(nullptr));
    for (void* p : {(void*)d_leaf, (void*)d_par, (void*)d_loc, (void*)d_dx, (void*)d_dy, (void*)d_off, (void*)d_i, (void*)d_j, (void*)d_dir, (void*)d_dof, (void*)d_sol, (void*)d_x, (void*)d_y})
        dev_free(p);
    if (e != cudaSuccess) return FEM2D_ERR_CUDA;
    std::copy(leaves.begin(), leaves.end(), leaf_ids);
    return FEM2D_OK;
}
