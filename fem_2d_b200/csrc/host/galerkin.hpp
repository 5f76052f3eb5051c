// C++ mirror of the reference's public surface for the assembly path, on top of the C-ABI (include/fem2d.h):
//   galerkin_sample_gep_hcurl<BSpace, AI, BI>(&domain, Option<[usize;2]>) -> Result<GEP, GalerkinSamplingError>
//   (src/fem_problem/galerkin.rs:33-40), GEP (linalg.rs:28-42), SparseMatrix (sparse_matrix.rs:12-17).
// The type parameters become tag types that map to the runtime kinds of the ABI.  The returned SparseMatrix keeps the
// reference's semantics (square symmetric, upper-triangular storage keyed [min,max] in (row, col) order, explicit zeros
// stored) but is backed by sorted arrays instead of a BTreeMap, which is what makes a 57 M-entry result affordable.
#pragma once
#include <array>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/fem2d.h"
#include "domain.hpp"
#include "glq.hpp"

namespace fem2d {

// ---- type tags standing in for the reference's generic arguments ----------------------------------------------------------
struct HierPoly { static constexpr int kind = FEM2D_BASIS_HIER_POLY; };            // hierarchical_basis_fns.rs:15
using KOLShapeFn = HierPoly;                                                        // pre-rename alias (README.md:10)
struct HierMaxOrtho { static constexpr int kind = FEM2D_BASIS_HIER_MAX_ORTHO; };   // hierarchical_basis_fns.rs:260
struct CurlCurl { static constexpr int kind = FEM2D_INTEGRAL_CURL_CURL; };         // integrals.rs:13
struct L2Inner { static constexpr int kind = FEM2D_INTEGRAL_L2_INNER; };           // integrals.rs:279

constexpr size_t MIN_GLQ_ORDER = 4;   // galerkin.rs:13

struct GalerkinSamplingError : std::runtime_error {   // galerkin.rs:191-213
    enum Kind { WrongContinuityCondition = FEM2D_ERR_WRONG_CONTINUITY, EmptyDOFSet = FEM2D_ERR_EMPTY_DOF_SET, InvalidGLQSettings = FEM2D_ERR_INVALID_GLQ } kind;
    explicit GalerkinSamplingError(Kind k) : std::runtime_error(fem2d_status_string((int)k)), kind(k) {}
};
struct BackendError : std::runtime_error {   // CUDA / argument failures of the native library (no reference counterpart)
    int status;
    BackendError(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

class SparseMatrix {   // sparse_matrix.rs:12-17
  public:
    size_t dimension = 0;
    std::vector<uint32_t> rows, cols;   // rows[k] <= cols[k], sorted by (row, col)
    std::vector<double> values;
    explicit SparseMatrix(size_t dim = 0) : dimension(dim) {
        if (dim > UINT32_MAX) throw std::runtime_error("Matrix Dimension cannot exceed the size of a u32!");   // sparse_matrix.rs:21-24
    }
    size_t num_entries() const {   // sparse_matrix.rs:32-35
        size_t diag = 0;
        for (size_t k = 0; k < rows.size(); k++) diag += rows[k] == cols[k];
        return 2 * rows.size() - diag;
    }
    template <class F> void iter_upper_tri(F&& f) const { for (size_t k = 0; k < rows.size(); k++) f(rows[k], cols[k], values[k]); }   // :123-127
    std::vector<double> to_dense() const {   // From<SparseMatrix> for DMatrix (:168-182), row-major
        std::vector<double> m(dimension * dimension, 0.0);
        for (size_t k = 0; k < rows.size(); k++) { m[rows[k] * dimension + cols[k]] = values[k]; m[cols[k] * dimension + rows[k]] = values[k]; }
        return m;
    }
};

struct GEP {   // linalg.rs:28-42
    SparseMatrix a, b;
    explicit GEP(size_t n = 0) : a(n), b(n) {}
};

// galerkin_sample_gep_hcurl (galerkin.rs:33-187).  `device` and `mode` are the only additions (defaults: GPU 0, bit-faithful).
template <class BSpace, class AI, class BI>
GEP galerkin_sample_gep_hcurl(const Domain& domain, std::optional<std::array<size_t, 2>> glq_grid_dim, int device = 0, int mode = FEM2D_MODE_EXACT) {
    if (domain.cc != ContinuityCondition::HCurl) throw GalerkinSamplingError(GalerkinSamplingError::WrongContinuityCondition);
    if (domain.dofs.empty()) throw GalerkinSamplingError(GalerkinSamplingError::EmptyDOFSet);
    const auto mo = domain.mesh.max_expansion_orders();
    size_t nu, nv;
    if (glq_grid_dim) {
        if ((*glq_grid_dim)[0] < MIN_GLQ_ORDER || (*glq_grid_dim)[1] < MIN_GLQ_ORDER) throw GalerkinSamplingError(GalerkinSamplingError::InvalidGLQSettings);
        nu = (*glq_grid_dim)[0]; nv = (*glq_grid_dim)[1];
    } else { nu = default_ngq(mo[0]); nv = default_ngq(mo[1]); }   // basis.rs:83-90
    std::vector<double> up, uw, vp, vw;
    gauss_quadrature_points(nu, false, up, uw);
    gauss_quadrature_points(nv, false, vp, vw);
    DomainView dv(domain);
    fem2d_plan* plan = nullptr;
    int st = fem2d_symbolic(&dv.view, device, 1, &plan);
    if (st >= FEM2D_ERR_WRONG_CONTINUITY && st <= FEM2D_ERR_INVALID_GLQ) throw GalerkinSamplingError((GalerkinSamplingError::Kind)st);
    if (st != FEM2D_OK) throw BackendError(st, fem2d_last_error());
    uint64_t info[16];
    fem2d_plan_info(plan, info);
    GEP gep(domain.dofs.size());
    gep.a.rows.resize(info[0]); gep.a.cols.resize(info[0]); gep.a.values.resize(info[0]); gep.b.values.resize(info[0]);
    st = fem2d_assemble(plan, BSpace::kind, AI::kind, BI::kind, mode, up.data(), uw.data(), (uint32_t)nu, vp.data(), vw.data(), (uint32_t)nv,
                        gep.a.rows.data(), gep.a.cols.data(), gep.a.values.data(), gep.b.values.data());
    fem2d_plan_free(plan);
    if (st >= FEM2D_ERR_WRONG_CONTINUITY && st <= FEM2D_ERR_INVALID_GLQ) throw GalerkinSamplingError((GalerkinSamplingError::Kind)st);
    if (st != FEM2D_OK) throw BackendError(st, fem2d_last_error());
    gep.b.rows = gep.a.rows; gep.b.cols = gep.a.cols;
    return gep;
}

}  // namespace fem2d
