// C-ABI of the host mirror (include/fem2d_host.h).  Pure host code.
#include "../../../include/fem2d_host.h"

#include <map>
#include <chrono>
#include <array>
#include <cstdio>
#include <memory>

#include "domain.hpp"
#include "glq.hpp"
#include "mesh_json.hpp"

struct fem2dh_mesh { fem2d::Mesh m; };
struct fem2dh_domain {
    fem2d::Domain d;
    std::unique_ptr<fem2d::DomainView> view;
};

namespace {
thread_local std::string g_herr;
int map_kind(fem2d::MeshError::Kind k) {
    using K = fem2d::MeshError;
    switch (k) {
        case K::ElemDoesNotExist: return FEM2DH_ERR_ELEM_DOES_NOT_EXIST;
        case K::ElemNotRefineable: return FEM2DH_ERR_ELEM_NOT_REFINEABLE;
        case K::DuplicateElemIds: return FEM2DH_ERR_DUPLICATE_ELEM_IDS;
        case K::ElemHasChildren: return FEM2DH_ERR_ELEM_HAS_CHILDREN;
        case K::EdgeHasChildren: return FEM2DH_ERR_EDGE_HAS_CHILDREN;
        case K::MinEdgeLength: return FEM2DH_ERR_MIN_EDGE_LENGTH;
        case K::EdgeOnEqualPoints: return FEM2DH_ERR_EDGE_ON_EQUAL_POINTS;
        case K::BisectionIdxExceeded: return FEM2DH_ERR_BISECTION_IDX_EXCEEDED;
        case K::RefinementOutOfBounds: return FEM2DH_ERR_REFINEMENT_OUT_OF_BOUNDS;
        case K::ExceededMaxExpansion: return FEM2DH_ERR_EXCEEDED_MAX_EXPANSION;
        case K::NegExpansion: return FEM2DH_ERR_NEG_EXPANSION;
        case K::BadMeshFile: return FEM2DH_ERR_BAD_MESH_FILE;
        default: return FEM2DH_ERR_INTERNAL;
    }
}
template <class F>
int guard(F&& f) {
    try { f(); return FEM2DH_OK; }
    catch (fem2d::MeshError& e) { g_herr = e.what(); return map_kind(e.kind); }
    catch (std::exception& e) { g_herr = e.what(); return FEM2DH_ERR_INTERNAL; }
}
fem2d::HRef make_href(int kind, int ext) {
    if (kind == FEM2DH_HREF_T) return fem2d::HRef::t();
    if (kind == FEM2DH_HREF_U) return ext < 0 ? fem2d::HRef::u() : fem2d::HRef::u_extended(ext);
    if (kind == FEM2DH_HREF_V) return ext < 0 ? fem2d::HRef::v() : fem2d::HRef::v_extended(ext);
    throw fem2d::MeshError(fem2d::MeshError::Internal, 0, "unknown HRef kind");
}
// Mutations are applied to a copy first so that a failing batch leaves the mesh untouched (mesh.rs:818, :1458).
template <class F>
int mutate(fem2dh_mesh* m, F&& f) {
    return guard([&] { fem2d::Mesh copy = m->m; f(copy); m->m = std::move(copy); });
}
}  // namespace

extern "C" {

const char* fem2dh_last_error(void) { return g_herr.c_str(); }

int fem2dh_mesh_from_file(const char* path, fem2dh_mesh** out) {
    return guard([&] { *out = nullptr; auto* h = new fem2dh_mesh(); try { h->m = fem2d::Mesh::from_file(path); } catch (...) { delete h; throw; } *out = h; });
}
int fem2dh_mesh_from_arrays(uint64_t ne, const double* mats, const int64_t* nids, uint64_t nn, const double* xy, fem2dh_mesh** out) {
    return guard([&] { *out = nullptr; auto* h = new fem2dh_mesh(); try { h->m = fem2d::Mesh::from_arrays(ne, mats, nids, nn, xy); } catch (...) { delete h; throw; } *out = h; });
}
int fem2dh_mesh_unit(fem2dh_mesh** out) { return guard([&] { *out = new fem2dh_mesh{fem2d::Mesh::unit()}; }); }
int fem2dh_mesh_clone(const fem2dh_mesh* m, fem2dh_mesh** out) { return guard([&] { *out = new fem2dh_mesh{m->m}; }); }
void fem2dh_mesh_free(fem2dh_mesh* m) { delete m; }

uint64_t fem2dh_mesh_num_elems(const fem2dh_mesh* m) { return m->m.elems.size(); }
uint64_t fem2dh_mesh_num_edges(const fem2dh_mesh* m) { return m->m.edges.size(); }
uint64_t fem2dh_mesh_num_nodes(const fem2dh_mesh* m) { return m->m.nodes.size(); }
uint64_t fem2dh_mesh_num_elements(const fem2dh_mesh* m) { return m->m.elements.size(); }

int fem2dh_mesh_elem_info(const fem2dh_mesh* m, uint64_t id, int64_t out[16], int64_t children[4]) {
    return guard([&] {
        if (id >= m->m.elems.size()) throw fem2d::MeshError(fem2d::MeshError::ElemDoesNotExist, id, "Attempt to access non-existent elem");
        const fem2d::Elem& e = m->m.elems[id];
        for (int k = 0; k < 4; k++) { out[k] = e.nodes[k]; out[4 + k] = e.edges[k]; }
        out[8] = e.parent; out[9] = e.has_children(); out[10] = e.poly_orders.ni; out[11] = e.poly_orders.nj; out[12] = e.h_u; out[13] = e.h_v;
        out[14] = e.element; out[15] = e.n_children;
        for (int k = 0; k < e.n_children; k++) children[k] = e.first_child + k;
    });
}
int fem2dh_mesh_edge_info(const fem2dh_mesh* m, uint64_t id, int64_t out[10], double* length) {
    return guard([&] {
        if (id >= m->m.edges.size()) throw fem2d::MeshError(fem2d::MeshError::Internal, id, "Attempt to access non-existent edge");
        const fem2d::Edge& e = m->m.edges[id];
        out[0] = e.nodes[0]; out[1] = e.nodes[1]; out[2] = e.boundary; out[3] = e.dir; out[4] = e.parent; out[5] = e.child[0]; out[6] = e.child[1];
        out[7] = e.active[0]; out[8] = e.active[1]; out[9] = e.child_node;
        if (length) *length = e.length;
    });
}
int fem2dh_mesh_node_info(const fem2dh_mesh* m, uint64_t id, double xy[2], int* boundary) {
    return guard([&] {
        if (id >= m->m.nodes.size()) throw fem2d::MeshError(fem2d::MeshError::Internal, id, "Attempt to access non-existent node");
        xy[0] = m->m.nodes[id].coords.x; xy[1] = m->m.nodes[id].coords.y;
        if (boundary) *boundary = m->m.nodes[id].boundary;
    });
}
int fem2dh_mesh_elem_range(const fem2dh_mesh* m, uint64_t id, int64_t from_ancestor, double out[4]) {
    return guard([&] {
        if (id >= m->m.elems.size()) throw fem2d::MeshError(fem2d::MeshError::ElemDoesNotExist, id, "Attempt to access non-existent elem");
        if (from_ancestor < 0) { for (int k = 0; k < 4; k++) out[k] = m->m.elems[id].range[k]; }
        else { auto r = m->m.relative_parametric_range(id, (size_t)from_ancestor); for (int k = 0; k < 4; k++) out[k] = r[k]; }
    });
}
int64_t fem2dh_mesh_descendant_elems(const fem2dh_mesh* m, uint64_t id, int include, int64_t* out, uint64_t cap) {
    int64_t n = -1;
    guard([&] { auto v = m->m.descendant_elems(id, include != 0); n = (int64_t)v.size(); for (size_t k = 0; k < v.size() && k < cap; k++) out[k] = (int64_t)v[k]; });
    return n;
}
int64_t fem2dh_mesh_ancestor_elems(const fem2dh_mesh* m, uint64_t id, int include, int64_t* out, uint64_t cap) {
    int64_t n = -1;
    guard([&] { auto v = m->m.ancestor_elems(id, include != 0); n = (int64_t)v.size(); for (size_t k = 0; k < v.size() && k < cap; k++) out[k] = (int64_t)v[k]; });
    return n;
}
void fem2dh_mesh_max_expansion_orders(const fem2dh_mesh* m, uint32_t out[2]) { auto o = m->m.max_expansion_orders(); out[0] = o[0]; out[1] = o[1]; }
int fem2dh_mesh_elem_is_h_refineable(const fem2dh_mesh* m, uint64_t id) {
    int r = -1;
    guard([&] { r = m->m.elem_is_h_refineable(id) ? 1 : 0; });
    return r;
}

int fem2dh_mesh_global_h_refinement(fem2dh_mesh* m, int kind, int ext) { return mutate(m, [&](fem2d::Mesh& x) { x.global_h_refinement(make_href(kind, ext)); }); }
int fem2dh_mesh_h_refine_elems(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, int kind, int ext) {
    return mutate(m, [&](fem2d::Mesh& x) { x.h_refine_elems(std::vector<size_t>(ids, ids + n), make_href(kind, ext)); });
}
int fem2dh_mesh_execute_h_refinements(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, const int32_t* kinds, const int32_t* exts) {
    return mutate(m, [&](fem2d::Mesh& x) {
        std::vector<std::pair<size_t, fem2d::HRef>> r;
        for (uint64_t k = 0; k < n; k++) r.push_back({(size_t)ids[k], make_href(kinds[k], exts ? exts[k] : -1)});
        x.execute_h_refinements(r);
    });
}
int fem2dh_mesh_global_p_refinement(fem2dh_mesh* m, int di, int dj) { return mutate(m, [&](fem2d::Mesh& x) { x.global_p_refinement(fem2d::PRef::from(di, dj)); }); }
int fem2dh_mesh_p_refine_elems(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, int di, int dj) {
    return mutate(m, [&](fem2d::Mesh& x) { x.p_refine_elems(std::vector<size_t>(ids, ids + n), fem2d::PRef::from(di, dj)); });
}
int fem2dh_mesh_execute_p_refinements(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, const int32_t* di, const int32_t* dj) {
    return mutate(m, [&](fem2d::Mesh& x) {
        std::vector<std::pair<size_t, fem2d::PRef>> r;
        for (uint64_t k = 0; k < n; k++) r.push_back({(size_t)ids[k], fem2d::PRef::from(di[k], dj[k])});
        x.execute_p_refinements(r);
    });
}
int fem2dh_mesh_set_global_expansion_orders(fem2dh_mesh* m, int ni, int nj) {
    return mutate(m, [&](fem2d::Mesh& x) {
        if (ni < 0 || nj < 0 || ni > 255 || nj > 255) throw fem2d::MeshError(fem2d::MeshError::ExceededMaxExpansion, 0, "orders out of range");
        x.set_global_expansion_orders(fem2d::Orders::make(ni, nj));
    });
}
int fem2dh_mesh_set_expansion_orders(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, const int32_t* ni, const int32_t* nj) {
    return mutate(m, [&](fem2d::Mesh& x) {
        std::vector<std::pair<size_t, fem2d::Orders>> r;
        for (uint64_t k = 0; k < n; k++) {
            if (ni[k] < 0 || nj[k] < 0 || ni[k] > 255 || nj[k] > 255) throw fem2d::MeshError(fem2d::MeshError::ExceededMaxExpansion, ids[k], "orders out of range");
            r.push_back({(size_t)ids[k], fem2d::Orders::make(ni[k], nj[k])});
        }
        x.set_expansion_orders(r);
    });
}

int fem2dh_domain_from_mesh(const fem2dh_mesh* m, int continuity, fem2dh_domain** out) {
    return guard([&] {
        *out = nullptr;
        auto h = std::make_unique<fem2dh_domain>();
        h->d = fem2d::Domain::from_mesh(m->m, (fem2d::ContinuityCondition)continuity);
        *out = h.release();
    });
}
int fem2dh_domain_blank(int continuity, fem2dh_domain** out) {
    return guard([&] { auto h = std::make_unique<fem2dh_domain>(); h->d = fem2d::Domain::blank((fem2d::ContinuityCondition)continuity); *out = h.release(); });
}
void fem2dh_domain_free(fem2dh_domain* d) { delete d; }
const fem2dh_mesh* fem2dh_domain_mesh(const fem2dh_domain* d) {
    // fem2dh_mesh is a struct whose only member is a Mesh: the domain's mesh can be viewed through the same handle type.
    return reinterpret_cast<const fem2dh_mesh*>(&d->d.mesh);
}
uint64_t fem2dh_domain_num_dofs(const fem2dh_domain* d) { return d->d.dofs.size(); }
uint64_t fem2dh_domain_num_basis_specs(const fem2dh_domain* d, uint64_t e) { return e < d->d.basis_specs.size() ? d->d.basis_specs[e].size() : 0; }
int fem2dh_domain_basis_specs(const fem2dh_domain* d, uint64_t e, int32_t* i, int32_t* j, int32_t* dir, int64_t* dof) {
    return guard([&] {
        const auto& l = d->d.local_basis_specs(e);
        for (size_t k = 0; k < l.size(); k++) { i[k] = l[k].i; j[k] = l[k].j; dir[k] = (int)l[k].dir; dof[k] = l[k].dof_id; }
    });
}
const fem2d_domain_view* fem2dh_domain_view(fem2dh_domain* d) {
    if (!d->view) d->view = std::make_unique<fem2d::DomainView>(d->d);
    return &d->view->view;
}

int fem2dh_gauss_quadrature_points(uint32_t n, double* points, double* weights) {
    return guard([&] {
        std::vector<double> p, w;
        fem2d::gauss_quadrature_points(n, false, p, w);
        std::copy(p.begin(), p.end(), points); std::copy(w.begin(), w.end(), weights);
    });
}
uint64_t fem2dh_default_ngq(uint64_t max_order) { return fem2d::default_ngq(max_order); }

// AIJMatrixBinary (sparse_matrix.rs:184-264): header bytes 00 12 7B 50 ("\0{P" with the raw 0x12 inside the literal),
// then u32 BE rows, cols, nnz, per-row counts, column ids, f64 BE values; full symmetric rows sorted by (row, col).
int fem2dh_write_petsc_aij(const char* path, uint64_t dim, uint64_t nnz_upper, const uint32_t* rows, const uint32_t* cols, const double* values) {
    return guard([&] {
        std::vector<uint32_t> counts(dim, 0);
        for (uint64_t k = 0; k < nnz_upper; k++) { counts[rows[k]]++; if (rows[k] != cols[k]) counts[cols[k]]++; }
        std::vector<uint64_t> start(dim + 1, 0);
        for (uint64_t r = 0; r < dim; r++) start[r + 1] = start[r] + counts[r];
        const uint64_t nnz = start[dim];
        std::vector<uint32_t> j(nnz); std::vector<double> a(nnz);
        std::vector<uint64_t> fill(start.begin(), start.end() - 1);
        // lower-triangle mirror entries (c, r) with r < c come first within row c (smaller column), in ascending r because the
        // upper-triangular input is sorted by (r, c); then the row's own upper part in ascending c.
        for (uint64_t k = 0; k < nnz_upper; k++) if (rows[k] != cols[k]) { const uint64_t p = fill[cols[k]]++; j[p] = rows[k]; a[p] = values[k]; }
        for (uint64_t k = 0; k < nnz_upper; k++) { const uint64_t p = fill[rows[k]]++; j[p] = cols[k]; a[p] = values[k]; }
        FILE* f = std::fopen(path, "wb");
        if (!f) throw std::runtime_error(std::string("cannot open ") + path);
        auto put32 = [&](uint32_t v) { unsigned char b[4] = {(unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v}; std::fwrite(b, 1, 4, f); };
        put32(1211216u); put32((uint32_t)dim); put32((uint32_t)dim); put32((uint32_t)nnz);
        for (uint64_t r = 0; r < dim; r++) put32(counts[r]);
        for (uint64_t k = 0; k < nnz; k++) put32(j[k]);
        for (uint64_t k = 0; k < nnz; k++) {
            uint64_t bits; std::memcpy(&bits, &a[k], 8);
            unsigned char b[8]; for (int q = 0; q < 8; q++) b[q] = (unsigned char)(bits >> (56 - 8 * q));
            std::fwrite(b, 1, 8, f);
        }
        std::fclose(f);
    });
}

// Stand-in for the caller-side rebuild of the reference's container from the sorted output arrays (`SparseMatrix::from_sorted_upper_tri` of
// a BTreeMap<[u32; 2], f64> bulk build): an ordered std::map filled with end hints, i.e. the O(n) best case.  Returns seconds.
double fem2dh_ordered_map_rebuild_seconds(uint64_t nnz, const uint32_t* rows, const uint32_t* cols, const double* values) {
    const auto t0 = std::chrono::steady_clock::now();
    {
        std::map<std::array<uint32_t, 2>, double> m;
        for (uint64_t k = 0; k < nnz; k++) m.emplace_hint(m.end(), std::array<uint32_t, 2>{rows[k], cols[k]}, values[k]);
        if (m.size() != nnz) return -1.0;
    }   // the map's destruction belongs to its owner, not to the rebuild; it is left out of the timed interval below on purpose
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
