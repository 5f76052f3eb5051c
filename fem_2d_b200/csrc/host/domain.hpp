// Host-side mirror of the reference's `Domain` (src/fem_domain/domain.rs, domain/dof.rs, domain/dof/basis_spec.rs) and the
// flattening of a Domain into the C-ABI view of include/fem2d.h.
//
// DoF numbering is reproduced constructively instead of through the reference's generate-everything-then-match maps:
//   * Elem-type DoFs: leaf Elems in id order; U specs with j >= 2 (i-major), then V specs with i >= 2 (i-major)
//     (basis_spec.rs:49-63 location rule, p_refinement.rs:49-64 order, domain.rs:83-96).
//   * Edge-type DoFs: edges in id order that have an active Elem pair; the lower-id active Elem's specs on that edge in
//     generation order, each matched with the same-order spec across the edge (basis_spec.rs:87-93, domain.rs:99-149).
// Citations are relative to /root/reference/.
#pragma once
#include <memory>

#include "../../../include/fem2d.h"
#include "mesh.hpp"

namespace fem2d {

enum class ContinuityCondition : uint32_t { HCurl = FEM2D_CC_HCURL, HDiv = FEM2D_CC_HDIV, Discontinuous = FEM2D_CC_DISCONTINUOUS };  // domain.rs:19-23
enum class BasisDir : uint8_t { U = 0, V = 1, W = 2 };   // basis_spec.rs:178-185

struct BSAddress { size_t elem_id, elem_idx; };   // basis_spec.rs:222-227

struct BasisSpec {   // basis_spec.rs:9-26
    uint8_t i, j;
    BasisDir dir;
    uint32_t elem_id;
    uint32_t elem_idx;
    uint32_t dof_id;
    int8_t edge_slot;   // -1: Elem-type (BasisLoc::ElemBs); 0..3: BasisLoc::EdgeBs(slot, elem.edges[slot])
    // (orders, dir, dof) as consumed by the integrators (basis_spec.rs:138-149)
    std::tuple<std::array<size_t, 2>, BasisDir, size_t> integration_data() const { return {{i, j}, dir, dof_id}; }
};

struct DoF {   // dof.rs:11-14: 1 address (Elem-type) or 2 (edge-type)
    size_t id;
    std::vector<BSAddress> basis_specs;
    const std::vector<BSAddress>& get_basis_specs() const { return basis_specs; }
};

class Domain {
  public:
    Mesh mesh;
    std::vector<DoF> dofs;
    std::vector<std::vector<BasisSpec>> basis_specs;
    ContinuityCondition cc = ContinuityCondition::HCurl;

    static Domain blank(ContinuityCondition cc) { Domain d; d.cc = cc; return d; }   // domain.rs:54-61
    static Domain unit(ContinuityCondition cc) { return from_mesh(Mesh::unit(), cc); }

    static Domain from_mesh(Mesh mesh_in, ContinuityCondition cc) {   // domain.rs:69-159
        Domain d;
        d.mesh = std::move(mesh_in);
        d.cc = cc;
        Mesh& mesh = d.mesh;
        mesh.set_edge_activation();
        if (cc != ContinuityCondition::HCurl && !mesh.elems.empty())
            throw MeshError(MeshError::Internal, 0, "not implemented: only the H(Curl) continuity condition is supported (basis_spec.rs:62)");
        d.basis_specs.assign(mesh.elems.size(), {});
        uint32_t next_dof = 0;
        auto push = [&](uint32_t elem, uint8_t i, uint8_t j, BasisDir dir, int8_t slot, uint32_t dof) {
            auto& list = d.basis_specs[elem];
            list.push_back(BasisSpec{i, j, dir, elem, (uint32_t)list.size(), dof, slot});
            return BSAddress{elem, list.size() - 1};
        };
        for (const Elem& e : mesh.elems) {
            if (e.has_children()) continue;
            const int ni = e.poly_orders.ni, nj = e.poly_orders.nj;
            for (int i = 0; i < ni; i++) for (int j = 2; j <= nj; j++) {
                BSAddress a = push(e.id, (uint8_t)i, (uint8_t)j, BasisDir::U, -1, next_dof);
                d.dofs.push_back(DoF{next_dof++, {a}});
            }
            for (int i = 2; i <= ni; i++) for (int j = 0; j < nj; j++) {
                BSAddress a = push(e.id, (uint8_t)i, (uint8_t)j, BasisDir::V, -1, next_dof);
                d.dofs.push_back(DoF{next_dof++, {a}});
            }
        }
        for (const Edge& ed : mesh.edges) {
            if (!ed.has_active_pair()) continue;
            const uint32_t lo = (uint32_t)std::min(ed.active[0], ed.active[1]), hi = (uint32_t)std::max(ed.active[0], ed.active[1]);
            const Elem& a = mesh.elems[lo]; const Elem& b = mesh.elems[hi];
            int sa = -1, sb = -1;
            for (int k = 0; k < 4; k++) { if (a.edges[k] == ed.id) sa = k; if (b.edges[k] == ed.id) sb = k; }
            if (sa < 0 || sb < 0) throw MeshError(MeshError::Internal, ed.id, "active Elem does not reference its Edge");
            if (sa <= 1) {   // U-directed tangential functions: j = slot, same i on both sides, slots sum to 1
                if (sa + sb != 1) continue;
                const int n = std::min(a.poly_orders.ni, b.poly_orders.ni);
                for (int i = 0; i < n; i++) {
                    BSAddress x = push(lo, (uint8_t)i, (uint8_t)sa, BasisDir::U, (int8_t)sa, next_dof);
                    BSAddress y = push(hi, (uint8_t)i, (uint8_t)sb, BasisDir::U, (int8_t)sb, next_dof);
                    d.dofs.push_back(DoF{next_dof++, {x, y}});
                }
            } else {         // V-directed: i = slot - 2, same j on both sides, slots sum to 5
                if (sa + sb != 5) continue;
                const int n = std::min(a.poly_orders.nj, b.poly_orders.nj);
                for (int j = 0; j < n; j++) {
                    BSAddress x = push(lo, (uint8_t)(sa - 2), (uint8_t)j, BasisDir::V, (int8_t)sa, next_dof);
                    BSAddress y = push(hi, (uint8_t)(sb - 2), (uint8_t)j, BasisDir::V, (int8_t)sb, next_dof);
                    d.dofs.push_back(DoF{next_dof++, {x, y}});
                }
            }
        }
        return d;
    }

    const std::vector<BasisSpec>& local_basis_specs(size_t elem_id) const {   // domain.rs:253-259
        if (elem_id >= mesh.elems.size()) throw MeshError(MeshError::ElemDoesNotExist, elem_id, "Attempt to access non-existent elem");
        return basis_specs[elem_id];
    }
    std::vector<std::pair<size_t, const std::vector<BasisSpec>*>> descendant_basis_specs(size_t elem_id) const {   // domain.rs:291-304
        std::vector<std::pair<size_t, const std::vector<BasisSpec>*>> out;
        for (size_t d : mesh.descendant_elems(elem_id, false)) out.push_back({d, &basis_specs[d]});
        return out;
    }
    std::vector<std::pair<size_t, const std::vector<BasisSpec>*>> ancestor_basis_specs(size_t elem_id) const {     // domain.rs:339-352
        std::vector<std::pair<size_t, const std::vector<BasisSpec>*>> out;
        for (size_t a : mesh.ancestor_elems(elem_id, false)) out.push_back({a, &basis_specs[a]});
        return out;
    }
    const BasisSpec& get_basis_spec(BSAddress a) const {   // domain.rs:183-199 (bound check fixed: >=)
        if (a.elem_id >= mesh.elems.size() || a.elem_idx >= basis_specs[a.elem_id].size()) throw MeshError(MeshError::ElemDoesNotExist, a.elem_id, "no such BasisSpec");
        return basis_specs[a.elem_id][a.elem_idx];
    }
};

// Owning flat copy of a Domain in the layout of fem2d_domain_view.  This is what the Rust shim builds from `&Domain`.
struct DomainView {
    std::vector<uint32_t> elem_element, bs_off, bs_dof;
    std::vector<int32_t> elem_parent;
    std::vector<uint8_t> elem_loc, bs_i, bs_j, bs_dir;
    std::vector<double> element_p0, element_p3, eps, mu;
    fem2d_domain_view view{};

    explicit DomainView(const Domain& d) {
        const Mesh& m = d.mesh;
        const size_t ne = m.elems.size();
        elem_element.resize(ne); elem_parent.resize(ne); elem_loc.resize(ne); bs_off.assign(ne + 1, 0);
        for (size_t e = 0; e < ne; e++) {
            elem_element[e] = m.elems[e].element; elem_parent[e] = m.elems[e].parent; elem_loc[e] = m.elems[e].loc;
            bs_off[e + 1] = bs_off[e] + (uint32_t)d.basis_specs[e].size();
        }
        bs_i.reserve(bs_off[ne]); bs_j.reserve(bs_off[ne]); bs_dir.reserve(bs_off[ne]); bs_dof.reserve(bs_off[ne]);
        for (size_t e = 0; e < ne; e++)
            for (const BasisSpec& b : d.basis_specs[e]) { bs_i.push_back(b.i); bs_j.push_back(b.j); bs_dir.push_back((uint8_t)b.dir); bs_dof.push_back(b.dof_id); }
        for (const Element& el : m.elements) {
            element_p0.push_back(el.points[0].x); element_p0.push_back(el.points[0].y);
            element_p3.push_back(el.points[3].x); element_p3.push_back(el.points[3].y);
            eps.push_back(el.materials.eps_re); mu.push_back(el.materials.mu_re);
        }
        auto mo = m.max_expansion_orders();
        view.n_elems = (uint32_t)ne; view.n_elements = (uint32_t)m.elements.size(); view.n_dofs = (uint32_t)d.dofs.size();
        view.continuity = (uint32_t)d.cc;
        view.elem_element = elem_element.data(); view.elem_parent = elem_parent.data(); view.elem_loc = elem_loc.data();
        view.element_p0 = element_p0.data(); view.element_p3 = element_p3.data();
        view.element_eps_re = eps.data(); view.element_mu_re = mu.data();
        view.bs_off = bs_off.data(); view.bs_i = bs_i.data(); view.bs_j = bs_j.data(); view.bs_dir = bs_dir.data(); view.bs_dof = bs_dof.data();
        view.i_max = mo[0]; view.j_max = mo[1];
    }
    DomainView(const DomainView&) = delete;
    DomainView& operator=(const DomainView&) = delete;
};

}  // namespace fem2d
