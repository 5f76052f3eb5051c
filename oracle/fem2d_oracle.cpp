// =====================================================================================
// fem2d_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17) of the reference crate jeremiah-corrado/fem_2d for the path
//   galerkin_sample_gep_hcurl::<HierPoly|HierMaxOrtho, CurlCurl, L2Inner>(&domain, Some([i,j]))
// including everything needed to regenerate the *same DoF numbering* (Mesh loading,
// RBS h-refinement, edge activation, Domain::from_mesh), because no Rust toolchain exists
// in the build image.  It is the checker for the CUDA path; nothing in the product
// (fem_2d_b200/) may include, link or call it.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it.
//
// Parity pinning: this restatement is gated (tests/test_oracle_pinned.py) on the reference's
// own fixtures -- test_input/test_evec.dat + test_eval.dat (624 DoFs, residual ~1e-14, lib.rs:85-104),
// the nalgebra surrogate eigenvalue 2.6479657 (lib.rs:48-65), the 20-point GLQ table
// (glq.rs:255-343) and the structural doctests (mesh.rs, domain.rs, h_refinement.rs).
// GLQ *node bits* are unpinned (reference uses nalgebra 0.30.1 SymmetricEigen, not vendored;
// pinned to 1e-9 only by glq.rs:326-375) -> nodes/weights are inputs of the assembly.
// HierMaxOrtho: parity unpinned (feature-gated, uncompilable in the reference, no tests).
//
// Floating point: compile with -O2 -ffp-contract=off, no fast-math.  Every arithmetic
// expression on the path is written in the reference's evaluation order (file:line cited).
//
// All file:line citations are relative to /root/reference/.
// =====================================================================================
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace orc {

// ------------------------------------------------------------------ space.rs
// FloatRep (space.rs:342-382): rounded representation used for point ordering.
struct FloatRep {
    bool sign;       // is_sign_positive
    uint64_t bits;
    static FloatRep from(double value) {
        const double POINT_UNIQUENESS_ACCURACY = 1e-12;  // space.rs:214
        double a = std::fabs(value);
        double integer_part = std::trunc(a);
        double fract = a - std::trunc(a);
        double fractional_rounded = std::round(fract / POINT_UNIQUENESS_ACCURACY) * POINT_UNIQUENESS_ACCURACY;
        double total = integer_part + fractional_rounded;
        FloatRep r;
        r.sign = !std::signbit(value);
        std::memcpy(&r.bits, &total, 8);
        return r;
    }
    bool operator==(const FloatRep& o) const { return sign == o.sign && bits == o.bits; }
    bool operator!=(const FloatRep& o) const { return !(*this == o); }
};
// space.rs:373-381  (-1 less, 0 equal, +1 greater)
static int float_rep_cmp(const FloatRep& a, const FloatRep& b) {
    auto c = [](uint64_t x, uint64_t y) { return x < y ? -1 : (x > y ? 1 : 0); };
    if (a.sign && b.sign) return c(a.bits, b.bits);
    if (!a.sign && b.sign) return -1;
    if (a.sign && !b.sign) return 1;
    return -c(a.bits, b.bits);
}

enum class ParaDir { U, V };

struct Point {  // space.rs:218-283
    double x = 0, y = 0;
    FloatRep x_cmp{true, 0}, y_cmp{true, 0};
    Point() = default;
    Point(double x_, double y_) : x(x_), y(y_), x_cmp(FloatRep::from(x_)), y_cmp(FloatRep::from(y_)) {}
    static Point between(const Point& a, const Point& b) { return Point((a.x + b.x) / 2.0, (a.y + b.y) / 2.0); }
    ParaDir orientation_with(const Point& o) const {  // space.rs:250-267
        if (!(x_cmp != o.x_cmp || y_cmp != o.y_cmp)) throw std::runtime_error("orientation between identical points");
        double dx = std::fabs(o.x - x);
        double dy = std::fabs(o.y - y);
        double theta = std::atan(dy / dx);
        const double FRAC_PI_4 = 0.78539816339744830961566084581988;
        return theta < FRAC_PI_4 ? ParaDir::U : ParaDir::V;
    }
    double dist(const Point& o) const {  // space.rs:269-274
        double dx = std::fabs(o.x - x), dy = std::fabs(o.y - y);
        return std::sqrt(dx * dx + dy * dy);
    }
};

struct V2D {  // space.rs:10-115
    double a[2];
    double operator[](int i) const { return a[i]; }
    double dot_with(const V2D& o) const { return a[0] * o.a[0] + a[1] * o.a[1]; }  // :19-21
    static double dot(const V2D& p, const V2D& q) { return p.a[0] * q.a[0] + p.a[1] * q.a[1]; }  // :23-25
};
static inline V2D operator*(const V2D& v, double c) { return V2D{{v.a[0] * c, v.a[1] * c}}; }          // :90-97
static inline V2D operator*(const V2D& v, const V2D& o) { return V2D{{v.a[0] * o.a[0], v.a[1] * o.a[1]}}; }  // :99-115
static inline V2D operator/(const V2D& v, double d) { return V2D{{v.a[0] / d, v.a[1] / d}}; }          // :72-79

struct M2D {  // space.rs:124-165
    V2D u, v;
    double det() const { return u[0] * v[1] - u[1] * v[0]; }  // :138-140
    M2D inverse() const {                                       // :142-147
        M2D t{V2D{{v[1], -1.0 * u[1]}}, V2D{{-1.0 * v[0], u[0]}}};
        double d = det();
        return M2D{t.u / d, t.v / d};
    }
};

// ------------------------------------------------------------------ h_refinement.rs
enum class HKind : int { T = 0, U = 1, V = 2 };
struct HRef {
    HKind kind;
    int ext;  // -1 = None, 0 / 1 = Some(idx)
};
// h_refinement.rs:159-195 (AddAssign)
static void href_add_assign(HRef& self, const HRef& rhs) {
    switch (self.kind) {
        case HKind::T: return;
        case HKind::U:
            if (self.ext < 0) {
                if (rhs.kind == HKind::V || rhs.kind == HKind::T) self = HRef{HKind::T, -1};
            } else {
                if (rhs.kind == HKind::U && rhs.ext >= 0) { if (rhs.ext != self.ext) self = HRef{HKind::T, -1}; }
                else if (rhs.kind == HKind::V || rhs.kind == HKind::T) self = HRef{HKind::T, -1};
            }
            return;
        case HKind::V:
            if (self.ext < 0) {
                if (rhs.kind == HKind::U || rhs.kind == HKind::T) self = HRef{HKind::T, -1};
            } else {
                if (rhs.kind == HKind::V && rhs.ext >= 0) { if (rhs.ext != self.ext) self = HRef{HKind::T, -1}; }
                else if (rhs.kind == HKind::U || rhs.kind == HKind::T) self = HRef{HKind::T, -1};
            }
            return;
    }
}
enum HRefLoc : uint8_t { SW = 0, SE, NW, NE, W, E, S, N };
static HRefLoc href_loc(HKind k, int idx) {  // h_refinement.rs:136-156
    switch (k) {
        case HKind::T: return (HRefLoc)idx;            // SW,SE,NW,NE
        case HKind::U: return idx == 0 ? W : E;
        case HKind::V: return idx == 0 ? S : N;
    }
    return SW;
}
using Range2 = std::array<std::array<double, 2>, 2>;
// h_refinement.rs:247-279
static Range2 sub_range(HRefLoc loc, const Range2& r) {
    double min_u = r[0][0], max_u = r[0][1], min_v = r[1][0], max_v = r[1][1];
    double mid_u = (min_u + max_u) / 2.0, mid_v = (min_v + max_v) / 2.0;
    std::array<double, 2> ur, vr;
    switch (loc) {
        case SW: case NW: case W: ur = {min_u, mid_u}; break;
        case SE: case NE: case E: ur = {mid_u, max_u}; break;
        default: ur = {min_u, max_u};
    }
    switch (loc) {
        case SW: case SE: case S: vr = {min_v, mid_v}; break;
        case NW: case NE: case N: vr = {mid_v, max_v}; break;
        default: vr = {min_v, max_v};
    }
    return Range2{ur, vr};
}

// ------------------------------------------------------------------ element.rs
struct Element {  // element.rs:15-50
    size_t id;
    Point points[4];
    double eps_re, eps_im, mu_re, mu_im;
};
static double map_range(double val, double in_min, double in_max, double out_min, double out_max) {  // element.rs:76-78
    return (val - in_min) * (out_max - out_min) / (in_max - in_min) + out_min;
}
static M2D element_parametric_mapping(const Element& el, const Range2& r) {  // element.rs:33-50
    double u_min = r[0][0], u_max = r[0][1], v_min = r[1][0], v_max = r[1][1];
    double real_x_min = map_range(u_min, -1.0, 1.0, el.points[0].x, el.points[3].x);
    double real_x_max = map_range(u_max, -1.0, 1.0, el.points[0].x, el.points[3].x);
    double real_y_min = map_range(v_min, -1.0, 1.0, el.points[0].y, el.points[3].y);
    double real_y_max = map_range(v_max, -1.0, 1.0, el.points[0].y, el.points[3].y);
    double dx_du = (real_x_max - real_x_min) / 2.0;
    double dy_dv = (real_y_max - real_y_min) / 2.0;
    return M2D{V2D{{dx_du, 0.0}}, V2D{{0.0, dy_dv}}};
}
static int element_order_points(const Point& p0, const Point& p1) {  // element.rs:56-61
    return p0.orientation_with(p1) == ParaDir::U ? float_rep_cmp(p0.x_cmp, p1.x_cmp) : float_rep_cmp(p0.y_cmp, p1.y_cmp);
}

// ------------------------------------------------------------------ node.rs / edge.rs / elem.rs
struct Node { size_t id; Point coords; bool boundary; };

struct Elem {  // elem.rs:101-110
    size_t id;
    std::array<size_t, 4> nodes, edges;
    size_t element;  // index into Mesh::elements (Arc<Element>)
    uint8_t h_u = 0, h_v = 0;
    uint8_t ni = 1, nj = 1;
    bool has_children = false;
    std::vector<size_t> children;
    std::vector<std::pair<size_t, HRefLoc>> ancestors;
    std::optional<size_t> parent_id() const { return ancestors.empty() ? std::nullopt : std::optional<size_t>(ancestors.back().first); }
    Range2 parametric_range() const {  // elem.rs:191-197
        Range2 acc{{{-1.0, 1.0}, {-1.0, 1.0}}};
        for (auto& a : ancestors) acc = sub_range(a.second, acc);
        return acc;
    }
    Range2 relative_parametric_range(size_t from_ancestor) const {  // elem.rs:170-188
        size_t start = ancestors.size();
        for (size_t k = 0; k < ancestors.size(); k++) if (ancestors[k].first == from_ancestor) { start = k; break; }
        if (start == ancestors.size()) throw std::runtime_error("not an ancestor");
        Range2 acc{{{-1.0, 1.0}, {-1.0, 1.0}}};
        for (size_t k = start; k < ancestors.size(); k++) acc = sub_range(ancestors[k].second, acc);
        return acc;
    }
};
struct ElemUninit {  // elem.rs:244-355
    size_t id;
    std::array<std::optional<size_t>, 4> nodes, edges;
    size_t element;
    std::vector<std::pair<size_t, HRefLoc>> ancestors;
    uint8_t h_u, h_v, ni, nj;
    void set_node(size_t idx, size_t nid) {
        if (nodes[idx]) { if (*nodes[idx] != nid) throw std::runtime_error("node already set"); }
        else nodes[idx] = nid;
    }
    void set_edge(size_t idx, size_t eid) {
        if (edges[idx]) throw std::runtime_error("edge already set");
        edges[idx] = eid;
    }
};

struct Edge {  // edge.rs:65-76
    size_t id;
    std::array<size_t, 2> nodes;
    bool boundary;
    ParaDir dir;
    double length;
    std::optional<std::array<size_t, 2>> children;
    std::optional<size_t> parent;
    std::map<std::array<uint8_t, 2>, size_t> elems[2];
    std::optional<std::array<size_t, 2>> active_elems;
    std::optional<size_t> child_node;

    static Edge make(size_t id, const Node& n0, const Node& n1, bool boundary) {  // edge.rs:80-95
        Edge e;
        e.id = id; e.nodes = {n0.id, n1.id}; e.boundary = boundary;
        e.dir = n0.coords.orientation_with(n1.coords);
        e.length = n0.coords.dist(n1.coords);
        return e;
    }
    void connect_elem(const Elem& elem) {  // edge.rs:97-123
        int pos = -1;
        for (int k = 0; k < 4; k++) if (elem.edges[k] == id) { pos = k; break; }
        if (pos < 0) throw std::runtime_error("elem not connected to edge");
        // HLevels::edge_ranking, h_refinement.rs:31-36
        std::array<uint8_t, 2> address = dir == ParaDir::U ? std::array<uint8_t, 2>{elem.h_v, elem.h_u}
                                                           : std::array<uint8_t, 2>{elem.h_u, elem.h_v};
        int side = (pos == 0 || pos == 2) ? 1 : 0;
        auto it = elems[side].find(address);
        if (it != elems[side].end()) { if (it->second != elem.id) throw std::runtime_error("edge side already connected"); }
        elems[side][address] = elem.id;
    }
    std::optional<size_t> last_entry(int side) const {  // edge.rs:216-222
        if (elems[side].empty()) return std::nullopt;
        return elems[side].rbegin()->second;
    }
    bool set_activation() {  // edge.rs:203-214
        auto a = last_entry(0), b = last_entry(1);
        if (a && b) { active_elems = std::array<size_t, 2>{*a, *b}; return true; }
        active_elems = std::nullopt;
        return false;
    }
};

const double MIN_EDGE_LENGTH = 3.0518e-5;   // mesh.rs:36
const uint8_t MAX_POLYNOMIAL_ORDER = 20;    // mesh.rs:42

struct OrcError : std::runtime_error { using std::runtime_error::runtime_error; };

// ------------------------------------------------------------------ mesh.rs
struct Mesh {
    std::vector<std::shared_ptr<Element>> elements;
    std::vector<Elem> elems;
    std::vector<Node> nodes;
    std::vector<Edge> edges;

    // mesh.rs:59-88
    static Mesh unit() {
        Mesh m;
        Point pts[4] = {Point(-1.0, -1.0), Point(1.0, -1.0), Point(-1.0, 1.0), Point(1.0, 1.0)};
        auto el = std::make_shared<Element>();
        el->id = 0;
        for (int k = 0; k < 4; k++) el->points[k] = pts[k];
        el->eps_re = 1.0; el->eps_im = 0.0; el->mu_re = 1.0; el->mu_im = 0.0;  // Materials::default element.rs:98-105
        m.elements.push_back(el);
        for (size_t k = 0; k < 4; k++) m.nodes.push_back(Node{k, pts[k], true});
        m.edges.push_back(Edge::make(0, m.nodes[0], m.nodes[1], true));
        m.edges.push_back(Edge::make(1, m.nodes[2], m.nodes[3], true));
        m.edges.push_back(Edge::make(2, m.nodes[0], m.nodes[2], true));
        m.edges.push_back(Edge::make(3, m.nodes[1], m.nodes[3], true));
        Elem e;  // Elem::new(0, [0,1,2,3], [0,1,2,3], unit_element); note: not connected to its (boundary) edges
        e.id = 0; e.nodes = {0, 1, 2, 3}; e.edges = {0, 1, 2, 3}; e.element = 0;
        m.elems.push_back(e);
        return m;
    }

    // mesh.rs:142-327 (JSON already parsed by the caller into flat arrays)
    static Mesh from_arrays(int n_elements, const double* materials, const int64_t* node_ids, int n_nodes, const double* xy) {
        Mesh m;
        std::vector<Point> points;
        for (int i = 0; i < n_nodes; i++) points.emplace_back(xy[2 * i], xy[2 * i + 1]);
        for (int e = 0; e < n_elements; e++) {
            auto el = std::make_shared<Element>();
            el->id = e;
            for (int k = 0; k < 4; k++) el->points[k] = points[node_ids[4 * e + k]];
            el->eps_re = materials[4 * e + 0]; el->eps_im = materials[4 * e + 1];
            el->mu_re = materials[4 * e + 2]; el->mu_im = materials[4 * e + 3];
            m.elements.push_back(el);
        }
        std::vector<int> counts(n_nodes, 0);
        for (int e = 0; e < n_elements; e++) for (int k = 0; k < 4; k++) counts[node_ids[4 * e + k]]++;
        for (int i = 0; i < n_nodes; i++) {
            if (counts[i] > 4) throw OrcError("node shared by more than 4 elements");
            m.nodes.push_back(Node{(size_t)i, points[i], counts[i] < 4});
        }
        // mesh.rs:202-227 ; EDGE_IDX_DEFS mesh.rs:1680-1681
        const int EDGE_IDX_DEFS[4][3] = {{0, 1, 1}, {2, 3, 0}, {0, 2, 1}, {1, 3, 0}};
        std::map<std::array<size_t, 2>, std::array<std::optional<size_t>, 2>> edge_node_pairs;
        for (int e = 0; e < n_elements; e++) {
            for (auto& d : EDGE_IDX_DEFS) {
                std::array<size_t, 2> key{(size_t)node_ids[4 * e + d[0]], (size_t)node_ids[4 * e + d[1]]};
                auto it = edge_node_pairs.find(key);
                if (it != edge_node_pairs.end()) {
                    if (it->second[d[2]]) throw OrcError("edge side already set");
                    it->second[d[2]] = (size_t)e;
                } else {
                    std::array<std::optional<size_t>, 2> v{std::nullopt, std::nullopt};
                    v[d[2]] = (size_t)e;
                    edge_node_pairs[key] = v;
                }
            }
        }
        // mesh.rs:231-258
        size_t edge_id = 0;
        for (auto& kv : edge_node_pairs) {
            int cnt = (kv.second[0] ? 1 : 0) + (kv.second[1] ? 1 : 0);
            m.edges.push_back(Edge::make(edge_id, m.nodes[kv.first[0]], m.nodes[kv.first[1]], cnt == 1));
            edge_id++;
        }
        // mesh.rs:261-288
        std::vector<std::array<std::optional<size_t>, 4>> elem_edges(n_elements);
        edge_id = 0;
        for (auto& kv : edge_node_pairs) {
            for (int side = 0; side < 2; side++) {
                if (!kv.second[side]) continue;
                size_t elem_id = *kv.second[side];
                int edge_idx;
                ParaDir dir = m.edges[edge_id].dir;
                if (side == 0 && dir == ParaDir::U) edge_idx = 1;
                else if (side == 1 && dir == ParaDir::U) edge_idx = 0;
                else if (side == 0 && dir == ParaDir::V) edge_idx = 3;
                else edge_idx = 2;
                if (elem_edges[elem_id][edge_idx]) throw OrcError("elem edge already set");
                elem_edges[elem_id][edge_idx] = edge_id;
            }
            edge_id++;
        }
        // mesh.rs:291-315
        for (int e = 0; e < n_elements; e++) {
            Elem el;
            el.id = e;
            for (int k = 0; k < 4; k++) {
                el.nodes[k] = node_ids[4 * e + k];
                if (!elem_edges[e][k]) throw OrcError("elem missing an edge");
                el.edges[k] = *elem_edges[e][k];
            }
            el.element = e;
            for (auto eid : el.edges) m.edges[eid].connect_elem(el);
            m.elems.push_back(el);
        }
        m.set_edge_activation();
        return m;
    }

    // mesh.rs:470-493
    void rec_descendant_elems(size_t id, bool include, std::vector<size_t>& out) const {
        if (include) out.push_back(id);
        if (elems[id].has_children) for (auto c : elems[id].children) rec_descendant_elems(c, true, out);
    }
    std::vector<size_t> descendant_elems(size_t id, bool include) const {
        if (id >= elems.size()) throw OrcError("elem does not exist");
        std::vector<size_t> out; rec_descendant_elems(id, include, out); return out;
    }
    // mesh.rs:520-541
    std::vector<size_t> ancestor_elems(size_t id, bool include) const {
        if (id >= elems.size()) throw OrcError("elem does not exist");
        std::vector<size_t> out;
        if (include) out.push_back(id);
        size_t cur = id;
        while (auto p = elems[cur].parent_id()) { out.push_back(*p); cur = *p; }
        return out;
    }
    std::array<uint8_t, 2> max_expansion_orders() const {  // mesh.rs:626-630
        std::array<uint8_t, 2> acc{0, 0};
        for (auto& e : elems) { acc[0] = std::max(acc[0], e.ni); acc[1] = std::max(acc[1], e.nj); }
        return acc;
    }
    bool elem_is_h_refineable(size_t id) const {  // mesh.rs:641-653
        if (id >= elems.size()) throw OrcError("elem does not exist");
        const Elem& e = elems[id];
        if (e.has_children) return false;
        for (auto eid : e.edges) if (!(edges[eid].length > MIN_EDGE_LENGTH)) return false;
        return true;
    }

    // mesh.rs:844-914. Returns 0 ok / nonzero error code (no mutation on validation errors).
    void execute_h_refinements(const std::vector<std::pair<size_t, HRef>>& refinements) {
        std::map<size_t, HRef> refinements_map;
        for (auto& r : refinements) {
            if (r.first >= elems.size()) throw OrcError("ElemDoesNotExist");
            if (!elem_is_h_refineable(r.first)) throw OrcError("ElemNotRefineable");
            auto it = refinements_map.find(r.first);
            if (it != refinements_map.end()) href_add_assign(it->second, r.second);
            else refinements_map[r.first] = r.second;
        }
        std::vector<std::pair<size_t, HRef>> extensions;
        size_t elem_id_tracker = elems.size();
        size_t node_id_tracker = nodes.size();
        size_t edge_id_tracker = edges.size();
        for (auto& kv : refinements_map) {
            size_t elem_id = kv.first;
            HRef refinement = kv.second;
            std::vector<ElemUninit> uninit = elem_h_refine(elem_id, refinement, elem_id_tracker);
            std::vector<Elem> new_elems;
            switch (refinement.kind) {
                case HKind::T: new_elems = execute_t_refinement(uninit, elem_id, node_id_tracker, edge_id_tracker); break;
                case HKind::U:
                    new_elems = execute_u_refinement(uninit, elem_id, node_id_tracker, edge_id_tracker);
                    if (refinement.ext >= 0) extensions.push_back({new_elems[refinement.ext].id, HRef{HKind::V, -1}});
                    break;
                case HKind::V:
                    new_elems = execute_v_refinement(uninit, elem_id, node_id_tracker, edge_id_tracker);
                    if (refinement.ext >= 0) extensions.push_back({new_elems[refinement.ext].id, HRef{HKind::U, -1}});
                    break;
            }
            for (auto& ne : new_elems) elems.push_back(ne);
        }
        if (!extensions.empty()) execute_h_refinements(extensions);
        set_edge_activation();
    }

    // elem.rs:128-155 + ElemUninit::new elem.rs:255-276
    std::vector<ElemUninit> elem_h_refine(size_t elem_id, HRef refinement, size_t& id_counter) {
        Elem& self = elems[elem_id];
        if (self.has_children) throw OrcError("ElemHasChildren");
        int n_children = refinement.kind == HKind::T ? 4 : 2;
        size_t starting_id = id_counter;
        id_counter += n_children;  // h_refinement.rs:119-134
        std::vector<ElemUninit> children;
        for (int idx = 0; idx < n_children; idx++) {
            ElemUninit c;
            c.id = starting_id + idx;
            c.element = self.element;
            c.ancestors = self.ancestors;
            c.ancestors.push_back({self.id, href_loc(refinement.kind, idx)});
            // HLevels::refined h_refinement.rs:22-28
            c.h_u = self.h_u + ((refinement.kind == HKind::T || refinement.kind == HKind::U) ? 1 : 0);
            c.h_v = self.h_v + ((refinement.kind == HKind::T || refinement.kind == HKind::V) ? 1 : 0);
            c.ni = self.ni; c.nj = self.nj;
            children.push_back(c);
        }
        self.has_children = true;
        self.children.clear();
        for (auto& c : children) self.children.push_back(c.id);
        return children;
    }

    // mesh.rs:1095-1126
    std::pair<std::array<size_t, 2>, size_t> h_refine_edge_if_needed(size_t parent_edge_id, size_t& node_id_tracker, size_t& edge_id_tracker) {
        if (edges[parent_edge_id].children) return {*edges[parent_edge_id].children, *edges[parent_edge_id].child_node};
        std::array<size_t, 2> new_edge_ids{edge_id_tracker, edge_id_tracker + 1};
        edge_id_tracker += 2;
        size_t new_node_id = node_id_tracker++;
        // Edge::h_refine edge.rs:126-170
        {
            Edge& pe = edges[parent_edge_id];
            double child_len = pe.length / 2.0;
            if (child_len < MIN_EDGE_LENGTH) throw OrcError("MinEdgeLength");
            pe.children = new_edge_ids;
            pe.child_node = new_node_id;
            Edge c0, c1;
            c0.id = new_edge_ids[0]; c0.nodes = {pe.nodes[0], new_node_id}; c0.boundary = pe.boundary; c0.dir = pe.dir;
            c0.length = child_len; c0.parent = pe.id;
            c1.id = new_edge_ids[1]; c1.nodes = {new_node_id, pe.nodes[1]}; c1.boundary = pe.boundary; c1.dir = pe.dir;
            c1.length = child_len; c1.parent = pe.id;
            edges.push_back(c0);
            edges.push_back(c1);
        }
        const Edge& pe = edges[parent_edge_id];
        Point coords = Point::between(nodes[pe.nodes[0]].coords, nodes[pe.nodes[1]].coords);
        if (new_node_id != nodes.size()) throw OrcError("node id mismatch");
        nodes.push_back(Node{new_node_id, coords, pe.boundary});
        return {new_edge_ids, new_node_id};
    }
    // mesh.rs:1128-1156
    size_t new_edge_between_nodes(std::array<size_t, 2> node_ids, size_t& edge_id_tracker, size_t parent_elem_id) {
        if (node_ids[0] == node_ids[1]) throw OrcError("identical nodes");
        size_t new_edge_id = edge_id_tracker++;
        const Node& n0 = nodes[node_ids[0]];
        const Node& n1 = nodes[node_ids[1]];
        (void)parent_elem_id;
        int ord = element_order_points(n0.coords, n1.coords);
        if (ord == 0) throw OrcError("EdgeOnEqualPoints");
        Edge e = ord < 0 ? Edge::make(new_edge_id, n0, n1, false) : Edge::make(new_edge_id, n1, n0, false);
        edges.push_back(e);
        return new_edge_id;
    }
    // mesh.rs:1158-1178 + ElemUninit::into_elem elem.rs:323-355
    std::vector<Elem> upgrade_uninit_elems(std::vector<ElemUninit>& uninit) {
        std::vector<Elem> out;
        for (auto& u : uninit) {
            Elem e;
            e.id = u.id;
            for (int k = 0; k < 4; k++) {
                if (!u.nodes[k] || !u.edges[k]) throw OrcError("UninitializedElem");
                e.nodes[k] = *u.nodes[k]; e.edges[k] = *u.edges[k];
            }
            e.element = u.element; e.ancestors = u.ancestors;
            e.h_u = u.h_u; e.h_v = u.h_v; e.ni = u.ni; e.nj = u.nj;
            out.push_back(e);
        }
        for (auto& e : out) for (auto eid : e.edges) edges[eid].connect_elem(e);
        return out;
    }
    // mesh.rs:916-979
    std::vector<Elem> execute_t_refinement(std::vector<ElemUninit>& ne, size_t parent, size_t& nt, size_t& et) {
        assert(ne.size() == 4);
        const Point& p0 = nodes[elems[parent].nodes[0]].coords;
        const Point& p3 = nodes[elems[parent].nodes[3]].coords;
        size_t center = nt++;
        Point cp = Point::between(p0, p3);
        if (center != nodes.size()) throw OrcError("node id mismatch");
        nodes.push_back(Node{center, cp, false});
        for (size_t idx = 0; idx < 4; idx++) {
            ne[idx].set_node(3 - idx, center);
            ne[idx].set_node(idx, elems[parent].nodes[idx]);
        }
        const int TBL[4][7] = {  // (edge_index, adj_child[2], shared_node_idx[2], internal_edge_idx[2]) mesh.rs:944-949
            {0, 0, 1, 1, 0, 3, 2}, {1, 2, 3, 3, 2, 3, 2}, {2, 0, 2, 2, 0, 1, 0}, {3, 1, 3, 3, 1, 1, 0}};
        for (auto& t : TBL) {
            auto [child_edge_ids, shared] = h_refine_edge_if_needed(elems[parent].edges[t[0]], nt, et);
            ne[t[1]].set_edge(t[0], child_edge_ids[0]);
            ne[t[2]].set_edge(t[0], child_edge_ids[1]);
            ne[t[1]].set_node(t[3], shared);
            ne[t[2]].set_node(t[4], shared);
            size_t new_edge = new_edge_between_nodes({shared, center}, et, parent);
            ne[t[1]].set_edge(t[5], new_edge);
            ne[t[2]].set_edge(t[6], new_edge);
        }
        return upgrade_uninit_elems(ne);
    }
    // mesh.rs:981-1036
    std::vector<Elem> execute_u_refinement(std::vector<ElemUninit>& ne, size_t parent, size_t& nt, size_t& et) {
        assert(ne.size() == 2);
        std::array<size_t, 2> outer{0, 0};
        const int TBL[2][5] = {{0, 1, 0, 0, 1}, {1, 3, 2, 2, 3}};  // (edge_index, shared_node_idx[2], outer_node_idx[2])
        for (auto& t : TBL) {
            auto [ce, shared] = h_refine_edge_if_needed(elems[parent].edges[t[0]], nt, et);
            outer[t[0]] = shared;
            ne[0].set_edge(t[0], ce[0]);
            ne[1].set_edge(t[0], ce[1]);
            ne[0].set_node(t[1], shared);
            ne[1].set_node(t[2], shared);
            ne[0].set_node(t[3], elems[parent].nodes[t[3]]);
            ne[1].set_node(t[4], elems[parent].nodes[t[4]]);
        }
        size_t new_edge = new_edge_between_nodes(outer, et, parent);
        ne[0].set_edge(3, new_edge);
        ne[1].set_edge(2, new_edge);
        ne[0].set_edge(2, elems[parent].edges[2]);
        ne[1].set_edge(3, elems[parent].edges[3]);
        return upgrade_uninit_elems(ne);
    }
    // mesh.rs:1038-1093
    std::vector<Elem> execute_v_refinement(std::vector<ElemUninit>& ne, size_t parent, size_t& nt, size_t& et) {
        assert(ne.size() == 2);
        std::array<size_t, 2> outer{0, 0};
        const int TBL[2][5] = {{2, 2, 0, 0, 2}, {3, 3, 1, 1, 3}};
        for (auto& t : TBL) {
            auto [ce, shared] = h_refine_edge_if_needed(elems[parent].edges[t[0]], nt, et);
            outer[t[0] - 2] = shared;
            ne[0].set_edge(t[0], ce[0]);
            ne[1].set_edge(t[0], ce[1]);
            ne[0].set_node(t[1], shared);
            ne[1].set_node(t[2], shared);
            ne[0].set_node(t[3], elems[parent].nodes[t[3]]);
            ne[1].set_node(t[4], elems[parent].nodes[t[4]]);
        }
        size_t new_edge = new_edge_between_nodes(outer, et, parent);
        ne[0].set_edge(1, new_edge);
        ne[1].set_edge(0, new_edge);
        ne[0].set_edge(0, elems[parent].edges[0]);
        ne[1].set_edge(1, elems[parent].edges[1]);
        return upgrade_uninit_elems(ne);
    }

    // mesh.rs:1180-1215
    void set_edge_activation() {
        for (auto& e : edges) e.active_elems = std::nullopt;
        std::vector<size_t> base;
        for (auto& e : edges) if (!e.parent && !e.boundary) base.push_back(e.id);
        for (auto b : base) if (!rec_set_edge_activation_in_tree(b)) throw OrcError("no active edge pair");
    }
    bool rec_set_edge_activation_in_tree(size_t edge_id) {
        if (edges[edge_id].set_activation()) {
            if (edges[edge_id].children) {
                auto ch = *edges[edge_id].children;
                bool a = rec_set_edge_activation_in_tree(ch[0]);
                bool b = rec_set_edge_activation_in_tree(ch[1]);
                if (a && b) edges[edge_id].active_elems = std::nullopt;
                else if (a != b) throw OrcError("inconsistent child edge support");
            }
            return true;
        }
        return false;
    }

    // p-refinement: mesh.rs:1486-1514 (execute_p_refinements), 1646-1665 (set_expansion_orders),
    // p_refinement.rs:20-42,104-123 (bounds)
    void execute_p_refinements(const std::vector<std::array<int64_t, 3>>& refs) {  // (elem_id, di, dj)
        std::map<size_t, std::array<int, 2>> rm;
        for (auto& r : refs) {
            if ((size_t)r[0] >= elems.size()) throw OrcError("ElemDoesNotExist");
            auto it = rm.find(r[0]);
            if (it != rm.end()) { it->second[0] += (int)r[1]; it->second[1] += (int)r[2]; }
            else rm[r[0]] = {(int)r[1], (int)r[2]};
        }
        for (auto& kv : rm) {
            const Elem& e = elems[kv.first];
            int lo_u = -((int)e.ni - 1), hi_u = MAX_POLYNOMIAL_ORDER - e.ni;
            int lo_v = -((int)e.nj - 1), hi_v = MAX_POLYNOMIAL_ORDER - e.nj;
            if (kv.second[0] < lo_u || kv.second[0] > hi_u || kv.second[1] < lo_v || kv.second[1] > hi_v)
                throw OrcError("RefinementOutOfBounds");
        }
        for (auto& kv : rm) { elems[kv.first].ni += kv.second[0]; elems[kv.first].nj += kv.second[1]; }
    }
    void set_expansion_orders(const std::vector<std::array<int64_t, 3>>& orders) {
        std::map<size_t, std::array<int, 2>> om;
        for (auto& o : orders) {
            if ((size_t)o[0] >= elems.size()) throw OrcError("ElemDoesNotExist");
            if (!om.insert({(size_t)o[0], {(int)o[1], (int)o[2]}}).second) throw OrcError("DuplicateElemIds");
        }
        for (auto& kv : om) {
            int ni = kv.second[0], nj = kv.second[1];
            if (ni > MAX_POLYNOMIAL_ORDER || nj > MAX_POLYNOMIAL_ORDER) throw OrcError("ExceededMaxExpansion");
            if (ni < 1 || nj < 1) throw OrcError("NegExpansion");
            elems[kv.first].ni = ni; elems[kv.first].nj = nj;
        }
    }
};

// ------------------------------------------------------------------ basis_spec.rs / domain.rs
enum class BasisDir : uint8_t { U = 0, V = 1, W = 2 };
struct BasisLoc { int kind; uint8_t idx; size_t id; };  // kind 0 ElemBs, 1 EdgeBs, 2 NodeBs
struct BasisSpec {
    size_t id; uint8_t i, j; BasisDir dir; size_t elem_id;
    std::optional<size_t> elem_idx, dof_id;
    BasisLoc loc;
};
static BasisSpec basis_spec_new(size_t id, uint8_t i, uint8_t j, BasisDir dir, const Elem& elem) {  // basis_spec.rs:41-75
    BasisLoc loc;
    auto edge_bs = [&](uint8_t idx) { return BasisLoc{1, idx, elem.edges[idx]}; };
    auto node_bs = [&](uint8_t idx) { return BasisLoc{2, idx, elem.nodes[idx]}; };
    if (i >= 2 && j >= 2) loc = BasisLoc{0, 0, 0};
    else if (j <= 1 && dir == BasisDir::U) loc = edge_bs(j);
    else if (i <= 1 && dir == BasisDir::V) loc = edge_bs(i + 2);
    else if (dir == BasisDir::W) {
        bool il = i < 2, jl = j < 2;
        if (il && !jl) loc = edge_bs(i + 2);
        else if (!il && jl) loc = edge_bs(j);
        else if (il && jl) loc = node_bs(i + 2 * j);
        else loc = BasisLoc{0, 0, 0};
    } else loc = BasisLoc{0, 0, 0};
    return BasisSpec{id, i, j, dir, elem.id, std::nullopt, std::nullopt, loc};
}
static bool matches_with_edge(const BasisSpec& a, const BasisSpec& b) {  // basis_spec.rs:80-110
    if (a.loc.kind != 1 || b.loc.kind != 1) throw OrcError("non-edge basis spec");
    if (a.loc.id != b.loc.id) throw OrcError("different edges");
    int idx0 = a.loc.idx, idx1 = b.loc.idx;
    if (a.dir == BasisDir::U && b.dir == BasisDir::U) return a.i == b.i && a.j + b.j == 1 && idx0 + idx1 == 1;
    if (a.dir == BasisDir::V && b.dir == BasisDir::V) return a.j == b.j && a.i + b.i == 1 && idx0 + idx1 == 5;
    if (a.dir == BasisDir::W && b.dir == BasisDir::W) {
        if (a.i >= 2 && b.i >= 2) return a.i == b.i && a.j + b.j == 1 && idx0 + idx1 == 1;
        if (a.j >= 2 && b.j >= 2) return a.j == b.j && a.i + b.i == 1 && idx0 + idx1 == 5;
        return false;
    }
    return false;
}
// p_refinement.rs:49-64
static std::vector<std::array<uint8_t, 2>> permutations(uint8_t ni, uint8_t nj, BasisDir dir) {
    std::vector<std::array<uint8_t, 2>> out;
    int imax = dir == BasisDir::U ? ni : ni + 1;   // exclusive
    int jmax = dir == BasisDir::V ? nj : nj + 1;
    for (int i = 0; i < imax; i++) for (int j = 0; j < jmax; j++) out.push_back({(uint8_t)i, (uint8_t)j});
    return out;
}

struct Domain {
    Mesh mesh;
    size_t n_dofs = 0;
    std::vector<std::vector<BasisSpec>> basis_specs;
    int cc = 0;  // 0 HCurl, 1 HDiv, 2 Discontinuous

    // domain.rs:69-159
    static Domain from_mesh(Mesh mesh_in, int cc) {
        Domain d;
        d.mesh = std::move(mesh_in);
        d.cc = cc;
        Mesh& mesh = d.mesh;
        mesh.set_edge_activation();
        d.basis_specs.assign(mesh.elems.size(), {});
        if (cc != 0) throw OrcError("unimplemented continuity condition");  // basis_spec.rs:62
        // gen_basis_specs domain.rs:201-235
        std::map<size_t, std::vector<BasisSpec>> elem_bs, edge_bs, node_bs;
        size_t bs_id = 0;
        for (auto& elem : mesh.elems) {
            for (BasisDir dir : {BasisDir::U, BasisDir::V, BasisDir::W}) {
                for (auto ij : permutations(elem.ni, elem.nj, dir)) {
                    BasisSpec bs = basis_spec_new(bs_id++, ij[0], ij[1], dir, elem);
                    if (bs.loc.kind == 0) elem_bs[elem.id].push_back(bs);
                    else if (bs.loc.kind == 1) edge_bs[bs.loc.id].push_back(bs);
                    else node_bs[bs.loc.id].push_back(bs);
                }
            }
        }
        size_t dof_id = 0;
        auto push_basis_spec = [&](BasisSpec bs, size_t dof) {  // domain.rs:356-368
            size_t eid = bs.elem_id;
            bs.dof_id = dof; bs.elem_idx = d.basis_specs[eid].size();
            d.basis_specs[eid].push_back(bs);
        };
        for (auto& kv : elem_bs) {  // domain.rs:83-96
            if (!mesh.elems[kv.first].has_children) {
                for (auto& bs : kv.second) {
                    if (bs.dir == BasisDir::U || bs.dir == BasisDir::V) push_basis_spec(bs, dof_id++);
                }
            }
        }
        for (auto& kv : edge_bs) {  // domain.rs:99-149
            auto& active = mesh.edges[kv.first].active_elems;
            if (!active) continue;
            std::vector<BasisSpec> rel;
            for (auto& bs : kv.second)
                if ((bs.dir == BasisDir::U || bs.dir == BasisDir::V) && (bs.elem_id == (*active)[0] || bs.elem_id == (*active)[1]))
                    rel.push_back(bs);
            std::vector<std::array<size_t, 2>> active_pairs;
            for (size_t a = 0; a < rel.size(); a++)
                for (size_t b = a + 1; b < rel.size(); b++)
                    if (matches_with_edge(rel[a], rel[b])) { active_pairs.push_back({a, b}); break; }
            for (auto& pr : active_pairs) {
                size_t id = dof_id++;
                push_basis_spec(rel[pr[0]], id);
                push_basis_spec(rel[pr[1]], id);
            }
        }
        d.n_dofs = dof_id;
        return d;
    }
};

// ------------------------------------------------------------------ glq.rs
// gauss_quadrature_points (glq.rs:179-222): Golub-Welsch on the Jacobi matrix.  The reference
// delegates the symmetric eigen-decomposition to nalgebra 0.30.1 SymmetricEigen (Cargo.lock:146-147,
// not vendored); restated here with the classic implicit-shift QL iteration on the tridiagonal
// matrix (same published algorithm family; bit-level parity of nodes is unpinned, see header).
static void glq_points(int n, std::vector<double>& pts, std::vector<double>& wts) {
    std::vector<double> d(n, 0.0), e(n, 0.0);
    for (int i = 1; i < n; i++) e[i - 1] = 0.5 / std::sqrt(1.0 - std::pow(2.0 * i, -2));  // glq.rs:180-182
    std::vector<double> z(n, 0.0);
    z[0] = 1.0;  // first row of the eigenvector matrix
    // tqli specialised to track only the first row of eigenvectors
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) throw OrcError("glq: too many iterations");
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i], b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                    s = f / r; c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    f = z[i + 1];
                    z[i + 1] = s * z[i] + c * f;
                    z[i] = c * z[i] - s * f;
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
            }
        } while (m != l);
    }
    std::vector<std::pair<double, double>> xw(n);
    for (int i = 0; i < n; i++) xw[i] = {d[i], z[i] * z[i] * 2.0};  // glq.rs:196-207
    std::sort(xw.begin(), xw.end(), [](auto& a, auto& b) { return a.first < b.first; });
    pts.resize(n); wts.resize(n);
    for (int i = 0; i < n; i++) { pts[i] = xw[i].first; wts[i] = xw[i].second; }
}
// glq.rs:238-249
static std::pair<double, std::vector<double>> scale_gauss_quad_points(const std::vector<double>& points, double mn, double mx) {
    double scale_factor = (mx - mn) / 2.0;
    double offset = (mx + mn) / 2.0;
    std::vector<double> out(points.size());
    for (size_t k = 0; k < points.size(); k++) out[k] = points[k] * scale_factor + offset;
    return {scale_factor, out};
}
// basis.rs:172-177
static size_t default_ngq(size_t max_order) {
    float conv = (float)(max_order * 4);
    int conv_p2 = (int)std::ceil(std::log2(conv));
    return (size_t)std::lround(std::pow(2.0f, (float)conv_p2));
}

// ------------------------------------------------------------------ hierarchical_basis_fns.rs
struct BSpaceTables {  // norm / norm_d1 / tang / tang_d1 [order][point]
    std::vector<std::vector<double>> norm, norm_d1, tang, tang_d1;
};
// HierPoly::new_without_d2, hierarchical_basis_fns.rs:102-162
static BSpaceTables hier_poly(size_t n_max, const std::vector<double>& points) {
    BSpaceTables t;
    size_t np = points.size();
    auto& pows = t.norm; auto& pows_d1 = t.norm_d1; auto& polys = t.tang; auto& polys_d1 = t.tang_d1;
    for (size_t n = 0; n <= n_max; n++) {
        double n_ = (double)n;
        if (n == 0) {
            std::vector<double> p(np); for (size_t k = 0; k < np; k++) p[k] = 1.0 - points[k];
            polys.push_back(p); polys_d1.push_back(std::vector<double>(np, -1.0));
            pows.push_back(std::vector<double>(np, 1.0)); pows_d1.push_back(std::vector<double>(np, 0.0));
        } else if (n == 1) {
            std::vector<double> p(np); for (size_t k = 0; k < np; k++) p[k] = 1.0 + points[k];
            polys.push_back(p); polys_d1.push_back(std::vector<double>(np, 1.0));
            pows.push_back(points); pows_d1.push_back(std::vector<double>(np, 1.0));
        } else {
            std::vector<double> pw(np), pd(np), pl(np), pld(np);
            for (size_t k = 0; k < np; k++) pw[k] = pows[n - 1][k] * points[k];
            for (size_t k = 0; k < np; k++) pd[k] = n_ * pows[n - 1][k];
            pows.push_back(pw); pows_d1.push_back(pd);
            if (n % 2 == 0) {
                for (size_t k = 0; k < np; k++) pl[k] = pw[k] - 1.0;
                pld = pd;
            } else {
                for (size_t k = 0; k < np; k++) pl[k] = pw[k] - points[k];
                for (size_t k = 0; k < np; k++) pld[k] = pd[k] - 1.0;
            }
            polys.push_back(pl); polys_d1.push_back(pld);
        }
    }
    return t;
}
// HierMaxOrtho (hierarchical_basis_fns.rs:206-293, 316-352, 425-463, 593-621); tables verbatim incl. apparent typos.
static const double EUC_NORM_COEFFS[12] = {0.968246, 2.561738, 0.838525, 4.248161, 0.816397, 5.882766, 0.808509, 1.0, 1.0, 1.0, 1.0, 1.0};
static const int Q_NUMERATORS[12][14] = {
    {-1, 0, 1}, {0, -3, 0, 3}, {-1, 0, -5, 0, 6}, {0, -3, 0, -7, 0, 10}, {-1, 0, -5, 0, -9, 0, 15},
    {0, -3, 0, -7, 0, -11, 0, 21}, {-1, 0, -5, 0, -9, 0, -13, 0, 28}, {0, -3, 0, -7, 0, -11, 0, -15, 0, 36},
    {-1, 0, -5, 0, -9, 0, -13, 0, -17, 0, 40}, {0, -3, 0, -7, 0, -11, 0, -15, 0, -19, 0, 55},
    {-1, 0, -5, 0, -9, 0, -13, 0, -17, 0, -21, 0, 66}, {0, -3, 0, -7, 0, -11, 0, -15, 0, -19, 0, -23, 0, 72}};
static const int Q_DENOMINATORS[12] = {1, 3, 6, 10, 15, 21, 28, 36, 40, 55, 66, 72};
static BSpaceTables hier_max_ortho(size_t n_max, const std::vector<double>& points) {
    if (n_max > 12) throw OrcError("HierMaxOrtho supports orders <= 12");  // Q_WEIGHTS has 11 rows (:240-252)
    BSpaceTables t;
    size_t np = points.size();
    auto& L = t.norm; auto& Ld = t.norm_d1;
    for (size_t i = 0; i <= n_max; i++) {  // LegendrePoly::with_specs_and_no_2nd_derivs :425-463
        L.push_back({}); Ld.push_back({});
        double i_f = (double)i;
        for (size_t p = 0; p < np; p++) {
            double point = points[p];
            if (i == 0) { L[i].push_back(1.0); Ld[i].push_back(0.0); }
            else if (i == 1) { L[i].push_back(point); Ld[i].push_back(1.0); }
            else {
                double v = ((2.0 * i_f - 1.0) * point * L[i - 1][p] - (i_f - 1.0) * L[i - 2][p]) / i_f;
                L[i].push_back(v);
                double pr = i_f * L[i - 1][p] + point * Ld[i - 1][p];
                Ld[i].push_back(pr);
            }
        }
    }
    for (size_t i = 0; i <= n_max; i++) {  // QFunction::with_specs_and_no_2nd_derivs :316-352
        if (i == 0) {
            std::vector<double> v(np); for (size_t p = 0; p < np; p++) v[p] = 1.0 - points[p];
            t.tang.push_back(v); t.tang_d1.push_back(std::vector<double>(np, -1.0));
        } else if (i == 1) {
            std::vector<double> v(np); for (size_t p = 0; p < np; p++) v[p] = 1.0 + points[p];
            t.tang.push_back(v); t.tang_d1.push_back(std::vector<double>(np, 1.0));
        } else {
            size_t dim = i + 1;  // get_q_weight_vector::<DIM>, index = DIM-3 = i-2 (:227-238)
            std::vector<double> w(dim);
            for (size_t k = 0; k < dim; k++) w[k] = ((double)Q_NUMERATORS[i - 2][k]) / ((double)Q_DENOMINATORS[i - 2]);
            double coeff = EUC_NORM_COEFFS[i - 2];
            std::vector<double> sv(np, 0.0), sp(np, 0.0);  // weighted_value_sum / weighted_prime_sum :593-621
            for (size_t order = 0; order < dim; order++) for (size_t p = 0; p < np; p++) sv[p] += w[order] * L[order][p];
            for (auto& s : sv) s *= coeff;
            for (size_t order = 0; order < dim; order++) for (size_t p = 0; p < np; p++) sp[p] += w[order] * Ld[order][p];
            for (auto& s : sp) s *= coeff;
            t.tang.push_back(sv); t.tang_d1.push_back(sp);
        }
    }
    return t;
}

// ------------------------------------------------------------------ basis.rs
struct HierCurlBasisFn {  // basis.rs:210-221
    std::vector<std::vector<M2D>> jac, jac_inv;
    std::vector<std::vector<double>> det_jac;
    V2D para_scale;
    BSpaceTables u_shapes, v_shapes;
    V2D f_u(size_t i, size_t j, size_t m, size_t n) const {  // :225-227
        return jac_inv[m][n].u * u_shapes.norm[i][m] * v_shapes.tang[j][n];
    }
    V2D f_v(size_t i, size_t j, size_t m, size_t n) const {  // :230-232
        return jac_inv[m][n].v * u_shapes.tang[i][m] * v_shapes.norm[j][n];
    }
    V2D f_u_d1(size_t i, size_t j, size_t m, size_t n, const V2D& ps) const {  // :235-242
        return jac_inv[m][n].u * V2D{{u_shapes.norm[i][m] * v_shapes.tang_d1[j][n], u_shapes.norm_d1[i][m] * v_shapes.tang[j][n]}} * ps;
    }
    V2D f_v_d1(size_t i, size_t j, size_t m, size_t n, const V2D& ps) const {  // :245-252
        return jac_inv[m][n].v * V2D{{u_shapes.tang[i][m] * v_shapes.norm_d1[j][n], u_shapes.tang_d1[i][m] * v_shapes.norm[j][n]}} * ps;
    }
    double glq_scale() const { return para_scale[0] * para_scale[1]; }            // :296-298
    double sample_scale(size_t m, size_t n) const { return det_jac[m][n]; }       // :330-332
    double uv_ratio(size_t m, size_t n) const { return jac[m][n].u[0] / jac[m][n].v[1]; }  // :341-343
    double vu_ratio(size_t m, size_t n) const { return jac[m][n].v[1] / jac[m][n].u[0]; }  // :346-348
};
// HierCurlBasisFn::defined_over basis.rs:365-423
static HierCurlBasisFn defined_over(const Mesh& mesh, const Elem& elem, const Elem* desc, const std::vector<double>& u_points,
                                    const std::vector<double>& v_points, size_t i_max, size_t j_max, int basis_kind) {
    double us = 1.0, vs = 1.0;
    std::vector<double> up = u_points, vp = v_points;
    if (desc && desc->id != elem.id) {
        Range2 r = desc->relative_parametric_range(elem.id);
        auto a = scale_gauss_quad_points(u_points, r[0][0], r[0][1]);
        auto b = scale_gauss_quad_points(v_points, r[1][0], r[1][1]);
        us = a.first; up = a.second; vs = b.first; vp = b.second;
    }
    HierCurlBasisFn f;
    const Element& el = *mesh.elements[elem.element];
    Range2 pr = elem.parametric_range();
    f.jac.resize(up.size()); f.jac_inv.resize(up.size()); f.det_jac.resize(up.size());
    for (size_t m = 0; m < up.size(); m++) {
        for (size_t n = 0; n < vp.size(); n++) {
            M2D t = element_parametric_mapping(el, pr);
            f.jac[m].push_back(t);
            f.jac_inv[m].push_back(t.inverse());
            f.det_jac[m].push_back(t.det());
        }
    }
    f.para_scale = V2D{{us, vs}};
    f.u_shapes = basis_kind == 0 ? hier_poly(i_max, up) : hier_max_ortho(i_max, up);
    f.v_shapes = basis_kind == 0 ? hier_poly(j_max, vp) : hier_max_ortho(j_max, vp);
    return f;
}

// ------------------------------------------------------------------ glq.rs:19-32 / integrals.rs
template <class F>
static double real_gauss_quad(const std::vector<double>& u_weights, const std::vector<double>& v_weights, F integrand) {
    double solution = 0.0;
    for (size_t m = 0; m < u_weights.size(); m++) {
        double inner_solution = 0.0;
        for (size_t n = 0; n < v_weights.size(); n++) inner_solution += integrand(m, n) * v_weights[n];
        solution += inner_solution * u_weights[m];
    }
    return solution;
}
static const V2D CURL_OP{{-1.0, 1.0}};  // integrals.rs:240
static double max_uv_ratios(const HierCurlBasisFn& p, const HierCurlBasisFn& q, size_t m, size_t n) {  // :250-259
    return (double)(uint8_t)(p.det_jac[m][n] >= q.det_jac[m][n]) * p.uv_ratio(m, n) +
           (double)(uint8_t)(p.det_jac[m][n] < q.det_jac[m][n]) * q.uv_ratio(m, n);
}
static double max_vu_ratios(const HierCurlBasisFn& p, const HierCurlBasisFn& q, size_t m, size_t n) {  // :262-271
    return (double)(uint8_t)(p.det_jac[m][n] >= q.det_jac[m][n]) * p.vu_ratio(m, n) +
           (double)(uint8_t)(p.det_jac[m][n] < q.det_jac[m][n]) * q.vu_ratio(m, n);
}
static double partial_max(double v1, double v2) { return v1 > v2 ? v1 : v2; }  // integrals.rs:421-423 (max_by: v2 on ties)

struct Integrator { std::vector<double> u_weights, v_weights; };

// CurlCurl::integrate integrals.rs:26-92
static double curl_curl_integrate(const Integrator& I, BasisDir pd, BasisDir qd, const std::array<size_t, 2>& po, const std::array<size_t, 2>& qo,
                                  const HierCurlBasisFn& P, const HierCurlBasisFn& Q, double mu_re) {
    double inner;
    if (pd == BasisDir::U && qd == BasisDir::U) {
        inner = real_gauss_quad(I.u_weights, I.v_weights, [&](size_t m, size_t n) {
            double p_curl = P.f_u_d1(po[0], po[1], m, n, Q.para_scale).dot_with(CURL_OP);
            double q_curl = Q.f_u_d1(qo[0], qo[1], m, n, P.para_scale).dot_with(CURL_OP);
            return p_curl * q_curl * max_uv_ratios(P, Q, m, n);
        });
    } else if (pd == BasisDir::U && qd == BasisDir::V) {
        inner = real_gauss_quad(I.u_weights, I.v_weights, [&](size_t m, size_t n) {
            double p_curl = P.f_u_d1(po[0], po[1], m, n, Q.para_scale).dot_with(CURL_OP);
            double q_curl = Q.f_v_d1(qo[0], qo[1], m, n, P.para_scale).dot_with(CURL_OP);
            return p_curl * q_curl;
        });
    } else if (pd == BasisDir::V && qd == BasisDir::U) {
        inner = real_gauss_quad(I.u_weights, I.v_weights, [&](size_t m, size_t n) {
            double p_curl = P.f_v_d1(po[0], po[1], m, n, Q.para_scale).dot_with(CURL_OP);
            double q_curl = Q.f_u_d1(qo[0], qo[1], m, n, P.para_scale).dot_with(CURL_OP);
            return p_curl * q_curl;
        });
    } else if (pd == BasisDir::V && qd == BasisDir::V) {
        inner = real_gauss_quad(I.u_weights, I.v_weights, [&](size_t m, size_t n) {
            double p_curl = P.f_v_d1(po[0], po[1], m, n, Q.para_scale).dot_with(CURL_OP);
            double q_curl = Q.f_v_d1(qo[0], qo[1], m, n, P.para_scale).dot_with(CURL_OP);
            return p_curl * q_curl * max_vu_ratios(P, Q, m, n);
        });
    } else inner = 0.0;
    return (1.0 / mu_re) * inner;
}
// L2Inner::integrate integrals.rs:292-354
static double l2_inner_integrate(const Integrator& I, BasisDir pd, BasisDir qd, const std::array<size_t, 2>& po, const std::array<size_t, 2>& qo,
                                 const HierCurlBasisFn& P, const HierCurlBasisFn& Q, double eps_re) {
    double inner;
    auto ev = [&](const HierCurlBasisFn& B, BasisDir d, const std::array<size_t, 2>& o, size_t m, size_t n) {
        return d == BasisDir::U ? B.f_u(o[0], o[1], m, n) : B.f_v(o[0], o[1], m, n);
    };
    if ((pd == BasisDir::U || pd == BasisDir::V) && (qd == BasisDir::U || qd == BasisDir::V)) {
        inner = real_gauss_quad(I.u_weights, I.v_weights, [&](size_t m, size_t n) {
            return V2D::dot(ev(P, pd, po, m, n), ev(Q, qd, qo, m, n)) * partial_max(P.sample_scale(m, n), Q.sample_scale(m, n));
        });
    } else inner = 0.0;
    return eps_re * P.glq_scale() * Q.glq_scale() * inner;
}

// ------------------------------------------------------------------ sparse_matrix.rs / linalg.rs / galerkin.rs
using Key = std::array<uint32_t, 2>;
struct SparseMatrix {  // sparse_matrix.rs:12-17
    size_t dimension;
    std::map<Key, double> entries;
    void insert_group(const std::vector<std::pair<std::array<size_t, 2>, double>>& g) {  // :68-98
        for (auto& e : g) {
            size_t r = e.first[0], c = e.first[1];
            if (r >= dimension || c >= dimension) throw OrcError("index exceeds dimension");
            Key k = r <= c ? Key{(uint32_t)r, (uint32_t)c} : Key{(uint32_t)c, (uint32_t)r};
            auto it = entries.find(k);
            if (it != entries.end()) it->second += e.second; else entries.emplace(k, e.second);
        }
    }
    void consume_matrix(SparseMatrix& other) {  // :106-120
        if (dimension != other.dimension) throw OrcError("dimension mismatch");
        std::map<Key, double> ne; ne.swap(other.entries);
        for (auto& kv : ne) {
            auto it = entries.find(kv.first);
            if (it != entries.end()) it->second += kv.second; else entries.emplace(kv.first, kv.second);
        }
    }
};

struct Sampler {  // BasisFnSampler basis.rs:53-134
    const Domain* domain;
    size_t i_max, j_max;
    int basis_kind;
    std::vector<double> u_points, v_points;
    std::mutex mtx;
    std::map<std::pair<size_t, int64_t>, std::shared_ptr<HierCurlBasisFn>> computed;
    std::shared_ptr<HierCurlBasisFn> sample_basis_fn(const Elem& elem, const Elem* desc) {
        std::pair<size_t, int64_t> key{elem.id, desc ? (int64_t)desc->id : -1};
        std::lock_guard<std::mutex> g(mtx);
        auto it = computed.find(key);
        if (it != computed.end()) return it->second;
        auto bs = std::make_shared<HierCurlBasisFn>(defined_over(domain->mesh, elem, desc, u_points, v_points, i_max, j_max, basis_kind));
        computed[key] = bs;
        return bs;
    }
};

struct GEPResult {
    int status = 0;  // 0 ok, 1 WrongContinuityCondition, 2 EmptyDOFSet, 3 InvalidGLQSettings  (galerkin.rs:191-195)
    std::vector<uint32_t> rows, cols;
    std::vector<double> a, b;
    double t_integrate = 0, t_merge = 0;
};
static const size_t MIN_GLQ_ORDER = 4;  // galerkin.rs:13

// one Elem's closure body, galerkin.rs:73-183
static void elem_matrices(const Domain& domain, Sampler& sampler, const Integrator& AI, const Integrator& BI, const Elem& elem,
                          SparseMatrix& local_a, SparseMatrix& local_b) {
    const Element& mats = *domain.mesh.elements[elem.element];
    auto bs_local = sampler.sample_basis_fn(elem, nullptr);
    const auto& local_basis_specs = domain.basis_specs[elem.id];
    std::vector<size_t> desc_ids = domain.mesh.descendant_elems(elem.id, false);
    std::vector<std::pair<std::array<size_t, 2>, double>> ea, eb;
    for (size_t i = 0; i < local_basis_specs.size(); i++) {  // local - local :91-127
        const BasisSpec& p = local_basis_specs[i];
        for (size_t k = i; k < local_basis_specs.size(); k++) {
            const BasisSpec& q = local_basis_specs[k];
            double a = curl_curl_integrate(AI, p.dir, q.dir, {p.i, p.j}, {q.i, q.j}, *bs_local, *bs_local, mats.mu_re);
            double b = l2_inner_integrate(BI, p.dir, q.dir, {p.i, p.j}, {q.i, q.j}, *bs_local, *bs_local, mats.eps_re);
            ea.push_back({{*p.dof_id, *q.dof_id}, a});
            eb.push_back({{*p.dof_id, *q.dof_id}, b});
        }
    }
    local_a.insert_group(ea);
    local_b.insert_group(eb);
    ea.clear(); eb.clear();
    for (const BasisSpec& p : local_basis_specs) {  // local - desc :138-178
        for (size_t q_elem_id : desc_ids) {
            const Elem& qe = domain.mesh.elems[q_elem_id];
            auto bs_p_sampled = sampler.sample_basis_fn(elem, &qe);
            auto bs_q_local = sampler.sample_basis_fn(qe, nullptr);
            for (const BasisSpec& q : domain.basis_specs[q_elem_id]) {
                double a = curl_curl_integrate(AI, p.dir, q.dir, {p.i, p.j}, {q.i, q.j}, *bs_p_sampled, *bs_q_local, mats.mu_re);
                double b = l2_inner_integrate(BI, p.dir, q.dir, {p.i, p.j}, {q.i, q.j}, *bs_p_sampled, *bs_q_local, mats.eps_re);
                ea.push_back({{*p.dof_id, *q.dof_id}, a});
                eb.push_back({{*p.dof_id, *q.dof_id}, b});
            }
        }
    }
    local_a.insert_group(ea);
    local_b.insert_group(eb);
}

// galerkin_sample_gep_hcurl galerkin.rs:33-187 + GEP::par_extend linalg.rs:59-81
static GEPResult galerkin_sample_gep_hcurl(const Domain& domain, int basis_kind, const double* u_pts, const double* u_w, size_t nu,
                                           const double* v_pts, const double* v_w, size_t nv, int n_threads) {
    GEPResult res;
    if (domain.cc != 0) { res.status = 1; return res; }
    if (domain.n_dofs == 0) { res.status = 2; return res; }
    if (nu < MIN_GLQ_ORDER || nv < MIN_GLQ_ORDER) { res.status = 3; return res; }
    auto mo = domain.mesh.max_expansion_orders();
    Sampler sampler;
    sampler.domain = &domain; sampler.i_max = mo[0]; sampler.j_max = mo[1]; sampler.basis_kind = basis_kind;
    sampler.u_points.assign(u_pts, u_pts + nu); sampler.v_points.assign(v_pts, v_pts + nv);
    Integrator AI{std::vector<double>(u_w, u_w + nu), std::vector<double>(v_w, v_w + nv)};
    Integrator BI = AI;
    SparseMatrix A{domain.n_dofs, {}}, B{domain.n_dofs, {}};
    auto t0 = std::chrono::steady_clock::now();
    size_t n_elems = domain.mesh.elems.size();
    std::deque<std::array<SparseMatrix, 2>> channel;  // unbounded mpsc channel: everything is buffered (linalg.rs:64-72)
    std::mutex ch_mtx;
    std::atomic<size_t> next{0};
    std::string err;
    auto worker = [&]() {
        try {
            for (;;) {
                size_t e = next.fetch_add(1);
                if (e >= n_elems) break;
                std::array<SparseMatrix, 2> lm{SparseMatrix{domain.n_dofs, {}}, SparseMatrix{domain.n_dofs, {}}};
                elem_matrices(domain, sampler, AI, BI, domain.mesh.elems[e], lm[0], lm[1]);
                std::lock_guard<std::mutex> g(ch_mtx);
                channel.push_back(std::move(lm));
            }
        } catch (std::exception& ex) { std::lock_guard<std::mutex> g(ch_mtx); err = ex.what(); }
    };
    if (n_threads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (!err.empty()) throw OrcError(err);
    auto t1 = std::chrono::steady_clock::now();
    for (auto& lm : channel) { A.consume_matrix(lm[0]); B.consume_matrix(lm[1]); }  // serial merge linalg.rs:74-79
    auto t2 = std::chrono::steady_clock::now();
    res.t_integrate = std::chrono::duration<double>(t1 - t0).count();
    res.t_merge = std::chrono::duration<double>(t2 - t1).count();
    if (A.entries.size() != B.entries.size()) throw OrcError("A/B key sets differ");
    res.rows.reserve(A.entries.size()); res.cols.reserve(A.entries.size());
    res.a.reserve(A.entries.size()); res.b.reserve(A.entries.size());
    auto ib = B.entries.begin();
    for (auto& kv : A.entries) {
        if (ib->first != kv.first) throw OrcError("A/B key sets differ");
        res.rows.push_back(kv.first[0]); res.cols.push_back(kv.first[1]);
        res.a.push_back(kv.second); res.b.push_back(ib->second);
        ++ib;
    }
    return res;
}

// Per-Elem matrices of selected Elems only (the closure body of galerkin.rs:73-183, not merged): used to spot-check
// full-size GPU results without running the whole CPU assembly.
struct ElemEntries { std::vector<int64_t> elem; std::vector<uint32_t> rows, cols; std::vector<double> a, b; };
static ElemEntries assemble_elems(const Domain& domain, int basis_kind, const double* u_pts, const double* u_w, size_t nu, const double* v_pts,
                                  const double* v_w, size_t nv, const int64_t* elem_ids, size_t n_ids) {
    auto mo = domain.mesh.max_expansion_orders();
    Sampler sampler;
    sampler.domain = &domain; sampler.i_max = mo[0]; sampler.j_max = mo[1]; sampler.basis_kind = basis_kind;
    sampler.u_points.assign(u_pts, u_pts + nu); sampler.v_points.assign(v_pts, v_pts + nv);
    Integrator AI{std::vector<double>(u_w, u_w + nu), std::vector<double>(v_w, v_w + nv)};
    Integrator BI = AI;
    ElemEntries out;
    for (size_t k = 0; k < n_ids; k++) {
        SparseMatrix la{domain.n_dofs, {}}, lb{domain.n_dofs, {}};
        elem_matrices(domain, sampler, AI, BI, domain.mesh.elems.at(elem_ids[k]), la, lb);
        auto ib = lb.entries.begin();
        for (auto& kv : la.entries) {
            out.elem.push_back(elem_ids[k]); out.rows.push_back(kv.first[0]); out.cols.push_back(kv.first[1]);
            out.a.push_back(kv.second); out.b.push_back(ib->second); ++ib;
        }
    }
    return out;
}

// UniformFieldSpace::xy_fields fields.rs:63-127 (+ uniform_range :407-410)
static void xy_fields(const Domain& domain, int basis_kind, size_t d0, size_t d1, const double* solution,
                      std::vector<int64_t>& leaf_ids, std::vector<double>& xv, std::vector<double>& yv) {
    auto uniform_range = [](double mn, double mx, size_t n) {
        double step = (mx - mn) / ((double)(n - 1));
        std::vector<double> r(n);
        for (size_t i = 0; i < n; i++) r[i] = ((double)i) * step + mn;
        return r;
    };
    std::vector<double> pu = uniform_range(-1.0, 1.0, d0), pv = uniform_range(-1.0, 1.0, d1);
    auto mo = domain.mesh.max_expansion_orders();
    for (auto& shell : domain.mesh.elems) {
        if (shell.has_children) continue;
        // reference allocates [d1][d0] but indexes [m<d0][n<d1] (quirk 5.8): only square densities are safe.
        if (d0 != d1) throw OrcError("xy_fields: non-square densities index out of bounds in the reference");
        std::vector<std::vector<double>> xs(d1, std::vector<double>(d0, 0.0)), ys(d1, std::vector<double>(d0, 0.0));
        for (size_t anc_id : domain.mesh.ancestor_elems(shell.id, true)) {
            HierCurlBasisFn bf = defined_over(domain.mesh, domain.mesh.elems[anc_id], &shell, pu, pv, mo[0], mo[1], basis_kind);
            for (auto& bs : domain.basis_specs[anc_id]) {
                for (size_t m = 0; m < d0; m++) for (size_t n = 0; n < d1; n++) {
                    V2D value = (bs.dir == BasisDir::U ? bf.f_u(bs.i, bs.j, m, n) : bs.dir == BasisDir::V ? bf.f_v(bs.i, bs.j, m, n) : V2D{{0.0, 0.0}}) *
                                solution[*bs.dof_id];
                    xs[m][n] += value[0];
                    ys[m][n] += value[1];
                }
            }
        }
        leaf_ids.push_back((int64_t)shell.id);
        for (auto& row : xs) for (double v : row) xv.push_back(v);
        for (auto& row : ys) for (double v : row) yv.push_back(v);
    }
}

}  // namespace orc

// =====================================================================================
// C interface for the Python test harness (ctypes).  Status: 0 ok, negative = exception
// (message via orc_last_error), positive = reference error enum.
// =====================================================================================
static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (std::exception & e) { g_err = e.what(); return ret; }

extern "C" {
const char* orc_last_error() { return g_err.c_str(); }

void* orc_mesh_from_arrays(int n_elements, const double* materials, const int64_t* node_ids, int n_nodes, const double* xy) {
    ORC_TRY return new orc::Mesh(orc::Mesh::from_arrays(n_elements, materials, node_ids, n_nodes, xy)); ORC_CATCH(nullptr)
}
void* orc_mesh_unit() { ORC_TRY return new orc::Mesh(orc::Mesh::unit()); ORC_CATCH(nullptr) }
void orc_mesh_free(void* m) { delete (orc::Mesh*)m; }
void* orc_mesh_clone(void* m) { return new orc::Mesh(*(orc::Mesh*)m); }
int64_t orc_mesh_num_elems(void* m) { return ((orc::Mesh*)m)->elems.size(); }
int64_t orc_mesh_num_edges(void* m) { return ((orc::Mesh*)m)->edges.size(); }
int64_t orc_mesh_num_nodes(void* m) { return ((orc::Mesh*)m)->nodes.size(); }
int orc_mesh_elem_is_h_refineable(void* m, int64_t id) { ORC_TRY return ((orc::Mesh*)m)->elem_is_h_refineable(id) ? 1 : 0; ORC_CATCH(-1) }
// kinds: 0 T, 1 U, 2 V ; ext: -1 none, 0/1
int orc_mesh_execute_h_refinements(void* m, int64_t n, const int64_t* ids, const int32_t* kinds, const int32_t* exts) {
    ORC_TRY
    std::vector<std::pair<size_t, orc::HRef>> r;
    for (int64_t k = 0; k < n; k++) r.push_back({(size_t)ids[k], orc::HRef{(orc::HKind)kinds[k], exts[k]}});
    // "If any errors are encountered, none of the refinements are executed" (mesh.rs:818): validate on a copy
    orc::Mesh copy = *(orc::Mesh*)m;
    copy.execute_h_refinements(r);
    *(orc::Mesh*)m = std::move(copy);
    return 0;
    ORC_CATCH(-1)
}
int orc_mesh_execute_p_refinements(void* m, int64_t n, const int64_t* ids, const int32_t* di, const int32_t* dj) {
    ORC_TRY
    std::vector<std::array<int64_t, 3>> r;
    for (int64_t k = 0; k < n; k++) r.push_back({ids[k], di[k], dj[k]});
    ((orc::Mesh*)m)->execute_p_refinements(r);
    return 0;
    ORC_CATCH(-1)
}
int orc_mesh_set_expansion_orders(void* m, int64_t n, const int64_t* ids, const int32_t* ni, const int32_t* nj) {
    ORC_TRY
    std::vector<std::array<int64_t, 3>> r;
    for (int64_t k = 0; k < n; k++) r.push_back({ids[k], ni[k], nj[k]});
    ((orc::Mesh*)m)->set_expansion_orders(r);
    return 0;
    ORC_CATCH(-1)
}
// per-Elem info: out[0..3] nodes, [4..7] edges, [8] parent (-1), [9] has_children, [10] ni, [11] nj, [12] h_u, [13] h_v, [14] element id, [15] n_children
int orc_mesh_elem_info(void* m, int64_t id, int64_t* out, int64_t* children4) {
    ORC_TRY
    auto& e = ((orc::Mesh*)m)->elems.at(id);
    for (int k = 0; k < 4; k++) { out[k] = e.nodes[k]; out[4 + k] = e.edges[k]; }
    auto p = e.parent_id();
    out[8] = p ? (int64_t)*p : -1; out[9] = e.has_children; out[10] = e.ni; out[11] = e.nj; out[12] = e.h_u; out[13] = e.h_v;
    out[14] = e.element; out[15] = e.children.size();
    for (size_t k = 0; k < e.children.size(); k++) children4[k] = e.children[k];
    return 0;
    ORC_CATCH(-1)
}
int orc_mesh_elem_ranges(void* m, int64_t id, int64_t from_ancestor, double* out4) {
    ORC_TRY
    auto& e = ((orc::Mesh*)m)->elems.at(id);
    orc::Range2 r = from_ancestor < 0 ? e.parametric_range() : e.relative_parametric_range(from_ancestor);
    out4[0] = r[0][0]; out4[1] = r[0][1]; out4[2] = r[1][0]; out4[3] = r[1][1];
    return 0;
    ORC_CATCH(-1)
}
// edge info: [0..1] nodes, [2] boundary, [3] dir (0 U,1 V), [4] parent(-1), [5..6] children(-1), [7..8] active pair (-1), [9] child node(-1)
int orc_mesh_edge_info(void* m, int64_t id, int64_t* out, double* length) {
    ORC_TRY
    auto& e = ((orc::Mesh*)m)->edges.at(id);
    out[0] = e.nodes[0]; out[1] = e.nodes[1]; out[2] = e.boundary; out[3] = e.dir == orc::ParaDir::U ? 0 : 1;
    out[4] = e.parent ? (int64_t)*e.parent : -1;
    out[5] = e.children ? (int64_t)(*e.children)[0] : -1; out[6] = e.children ? (int64_t)(*e.children)[1] : -1;
    out[7] = e.active_elems ? (int64_t)(*e.active_elems)[0] : -1; out[8] = e.active_elems ? (int64_t)(*e.active_elems)[1] : -1;
    out[9] = e.child_node ? (int64_t)*e.child_node : -1;
    *length = e.length;
    return 0;
    ORC_CATCH(-1)
}
int orc_mesh_node_xy(void* m, int64_t id, double* xy) {
    ORC_TRY auto& n = ((orc::Mesh*)m)->nodes.at(id); xy[0] = n.coords.x; xy[1] = n.coords.y; return n.boundary ? 1 : 0; ORC_CATCH(-1)
}
int64_t orc_mesh_descendant_elems(void* m, int64_t id, int include, int64_t* out, int64_t cap) {
    ORC_TRY
    auto v = ((orc::Mesh*)m)->descendant_elems(id, include != 0);
    for (size_t k = 0; k < v.size() && (int64_t)k < cap; k++) out[k] = v[k];
    return v.size();
    ORC_CATCH(-1)
}
int64_t orc_mesh_ancestor_elems(void* m, int64_t id, int include, int64_t* out, int64_t cap) {
    ORC_TRY
    auto v = ((orc::Mesh*)m)->ancestor_elems(id, include != 0);
    for (size_t k = 0; k < v.size() && (int64_t)k < cap; k++) out[k] = v[k];
    return v.size();
    ORC_CATCH(-1)
}
void orc_mesh_max_expansion_orders(void* m, int32_t* out2) {
    auto o = ((orc::Mesh*)m)->max_expansion_orders(); out2[0] = o[0]; out2[1] = o[1];
}
void orc_sub_range(int loc, const double* in4, double* out4) {
    orc::Range2 r{{{in4[0], in4[1]}, {in4[2], in4[3]}}};
    r = orc::sub_range((orc::HRefLoc)loc, r);
    out4[0] = r[0][0]; out4[1] = r[0][1]; out4[2] = r[1][0]; out4[3] = r[1][1];
}

void* orc_domain_from_mesh(void* m, int cc) {
    ORC_TRY return new orc::Domain(orc::Domain::from_mesh(*(orc::Mesh*)m, cc)); ORC_CATCH(nullptr)
}
void orc_domain_free(void* d) { delete (orc::Domain*)d; }
void* orc_domain_mesh(void* d) { return &((orc::Domain*)d)->mesh; }
int64_t orc_domain_num_dofs(void* d) { return ((orc::Domain*)d)->n_dofs; }
int64_t orc_domain_num_basis_specs(void* d, int64_t elem_id) { return ((orc::Domain*)d)->basis_specs.at(elem_id).size(); }
// basis specs of one Elem in reference list order: i, j, dir, dof
int orc_domain_basis_specs(void* d, int64_t elem_id, int32_t* i, int32_t* j, int32_t* dir, int64_t* dof) {
    ORC_TRY
    auto& l = ((orc::Domain*)d)->basis_specs.at(elem_id);
    for (size_t k = 0; k < l.size(); k++) { i[k] = l[k].i; j[k] = l[k].j; dir[k] = (int)l[k].dir; dof[k] = *l[k].dof_id; }
    return 0;
    ORC_CATCH(-1)
}

int orc_glq(int n, double* pts, double* wts) {
    ORC_TRY
    std::vector<double> p, w; orc::glq_points(n, p, w);
    std::copy(p.begin(), p.end(), pts); std::copy(w.begin(), w.end(), wts);
    return 0;
    ORC_CATCH(-1)
}
int64_t orc_default_ngq(int64_t max_order) { return orc::default_ngq(max_order); }
// basis tables: out arrays [(n_max+1) * n_points] each (norm, norm_d1, tang, tang_d1)
int orc_basis_tables(int basis_kind, int n_max, int np, const double* pts, double* norm, double* norm_d1, double* tang, double* tang_d1) {
    ORC_TRY
    std::vector<double> p(pts, pts + np);
    orc::BSpaceTables t = basis_kind == 0 ? orc::hier_poly(n_max, p) : orc::hier_max_ortho(n_max, p);
    for (int n = 0; n <= n_max; n++) for (int k = 0; k < np; k++) {
        norm[n * np + k] = t.norm[n][k]; norm_d1[n * np + k] = t.norm_d1[n][k];
        tang[n * np + k] = t.tang[n][k]; tang_d1[n * np + k] = t.tang_d1[n][k];
    }
    return 0;
    ORC_CATCH(-1)
}

void* orc_assemble(void* d, int basis_kind, const double* u_pts, const double* u_w, int64_t nu, const double* v_pts, const double* v_w,
                   int64_t nv, int n_threads) {
    ORC_TRY
    return new orc::GEPResult(orc::galerkin_sample_gep_hcurl(*(orc::Domain*)d, basis_kind, u_pts, u_w, nu, v_pts, v_w, nv, n_threads));
    ORC_CATCH(nullptr)
}
int orc_gep_status(void* g) { return ((orc::GEPResult*)g)->status; }
int64_t orc_gep_nnz(void* g) { return ((orc::GEPResult*)g)->rows.size(); }
void orc_gep_times(void* g, double* out2) { out2[0] = ((orc::GEPResult*)g)->t_integrate; out2[1] = ((orc::GEPResult*)g)->t_merge; }
void orc_gep_copy(void* g, uint32_t* rows, uint32_t* cols, double* a, double* b) {
    auto* r = (orc::GEPResult*)g;
    std::copy(r->rows.begin(), r->rows.end(), rows); std::copy(r->cols.begin(), r->cols.end(), cols);
    std::copy(r->a.begin(), r->a.end(), a); std::copy(r->b.begin(), r->b.end(), b);
}
void orc_gep_free(void* g) { delete (orc::GEPResult*)g; }

void* orc_assemble_elems(void* d, int basis_kind, const double* u_pts, const double* u_w, int64_t nu, const double* v_pts, const double* v_w,
                         int64_t nv, const int64_t* elem_ids, int64_t n_ids) {
    ORC_TRY
    return new orc::ElemEntries(orc::assemble_elems(*(orc::Domain*)d, basis_kind, u_pts, u_w, nu, v_pts, v_w, nv, elem_ids, n_ids));
    ORC_CATCH(nullptr)
}
int64_t orc_elem_entries_size(void* e) { return ((orc::ElemEntries*)e)->rows.size(); }
void orc_elem_entries_copy(void* e, int64_t* elem, uint32_t* rows, uint32_t* cols, double* a, double* b) {
    auto* r = (orc::ElemEntries*)e;
    std::copy(r->elem.begin(), r->elem.end(), elem); std::copy(r->rows.begin(), r->rows.end(), rows); std::copy(r->cols.begin(), r->cols.end(), cols);
    std::copy(r->a.begin(), r->a.end(), a); std::copy(r->b.begin(), r->b.end(), b);
}
void orc_elem_entries_free(void* e) { delete (orc::ElemEntries*)e; }

// xy_fields: returns number of leaves; outputs sized n_leaves * d0*d1 (caller passes capacity in leaves)
int64_t orc_xy_fields(void* d, int basis_kind, int64_t d0, int64_t d1, const double* solution, int64_t cap_leaves, int64_t* leaf_ids,
                      double* x_values, double* y_values) {
    ORC_TRY
    std::vector<int64_t> ids; std::vector<double> xv, yv;
    orc::xy_fields(*(orc::Domain*)d, basis_kind, d0, d1, solution, ids, xv, yv);
    if ((int64_t)ids.size() <= cap_leaves) {
        std::copy(ids.begin(), ids.end(), leaf_ids);
        std::copy(xv.begin(), xv.end(), x_values); std::copy(yv.begin(), yv.end(), y_values);
    }
    return ids.size();
    ORC_CATCH(-1)
}
}
