// Host-side mirror of the reference's `Mesh` (src/fem_domain/domain/mesh.rs and mesh/*.rs).
//
// The GPU path never mutates a mesh; this mirror exists so that C++/Python callers (tests, bench, users without
// the Rust crate) can build exactly the Domains the reference builds -- same Elem / Edge / Node ids, same edge
// activation -- and hand them to the C-ABI in include/fem2d.h.  It is written data-oriented (flat records, child
// ranges, cached dyadic parametric ranges, max-rank edge registries) rather than as the reference's pointer graph.
//
// Reference citations are relative to /root/reference/.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace fem2d {

constexpr double MIN_EDGE_LENGTH = 3.0518e-5;    // mesh.rs:36
constexpr uint8_t MAX_POLYNOMIAL_ORDER = 20;     // mesh.rs:42

// ---- errors (h_refinement.rs:283-296, p_refinement.rs PRefError, mesh.rs:1799-1804) -------------------------------
struct MeshError : std::runtime_error {
    enum Kind {
        ElemDoesNotExist, ElemNotRefineable, DuplicateElemIds, ElemHasChildren, EdgeHasChildren, MinEdgeLength,
        EdgeOnEqualPoints, BisectionIdxExceeded, RefinementOutOfBounds, ExceededMaxExpansion, NegExpansion, BadMeshFile,
        Internal
    } kind;
    size_t id;
    MeshError(Kind k, size_t id_, const std::string& what) : std::runtime_error(what), kind(k), id(id_) {}
};

// ---- HRef / PRef / Orders (h_refinement.rs:69-117, p_refinement.rs) -------------------------------------------------
struct HRef {
    enum Kind : uint8_t { T = 0, U = 1, V = 2 } kind = T;
    int8_t ext = -1;  // U(Some(k)) / V(Some(k)): follow-up bisection of child k in the other direction
    static HRef t() { return {T, -1}; }
    static HRef u() { return {U, -1}; }
    static HRef v() { return {V, -1}; }
    static HRef u_extended(int k) { check_ext(k); return {U, (int8_t)k}; }
    static HRef v_extended(int k) { check_ext(k); return {V, (int8_t)k}; }
    // Combination rule when one Elem receives several requests in a batch (h_refinement.rs:159-195): anything that asks for
    // both directions (or conflicting extensions) collapses to T.
    void merge(const HRef& o) {
        if (kind == T) return;
        bool other_dir = (o.kind == T) || (o.kind != kind);
        if (other_dir) { *this = t(); return; }
        if (ext >= 0 && o.ext >= 0 && o.ext != ext) *this = t();
    }
  private:
    static void check_ext(int k) {
        if (k != 0 && k != 1) throw MeshError(MeshError::BisectionIdxExceeded, 0, "Extended refinement index must be 0 or 1!");
    }
};
struct PRef { int8_t di = 0, dj = 0; static PRef from(int i, int j) { return {(int8_t)i, (int8_t)j}; } };
struct Orders { uint8_t ni = 1, nj = 1; static Orders make(int i, int j) { return {(uint8_t)i, (uint8_t)j}; } };

// HRefLoc codes shared with include/fem2d.h (h_refinement.rs:211-229)
enum Loc : uint8_t { SW = 0, SE, NW, NE, W, E, S, N, BASE = 255 };

// Child sub-range inside a parent range (h_refinement.rs:247-279); dyadic, hence exact.
inline void sub_range(uint8_t loc, const double in[4], double out[4]) {
    const double mu = (in[0] + in[1]) / 2.0, mv = (in[2] + in[3]) / 2.0;
    const bool west = (loc == SW || loc == NW || loc == W), east = (loc == SE || loc == NE || loc == E);
    const bool south = (loc == SW || loc == SE || loc == S), north = (loc == NW || loc == NE || loc == N);
    out[0] = east ? mu : in[0];
    out[1] = west ? mu : in[1];
    out[2] = north ? mv : in[2];
    out[3] = south ? mv : in[3];
}

// ---- geometry ---------------------------------------------------------------------------------------------------------
// Coordinates are compared through a representation rounded to 1e-12 (space.rs:342-382).
struct CoordKey {
    bool nonneg; uint64_t mag;
    explicit CoordKey(double v) {
        const double a = std::fabs(v), ip = std::trunc(a);
        const double r = ip + std::round((a - ip) / 1e-12) * 1e-12;
        nonneg = !std::signbit(v);
        std::memcpy(&mag, &r, 8);
    }
    int cmp(const CoordKey& o) const {
        if (nonneg != o.nonneg) return nonneg ? 1 : -1;
        int c = mag < o.mag ? -1 : (mag > o.mag ? 1 : 0);
        return nonneg ? c : -c;
    }
};
struct Point {
    double x = 0, y = 0;
    static Point mid(const Point& a, const Point& b) { return {(a.x + b.x) / 2.0, (a.y + b.y) / 2.0}; }  // space.rs:235
};
// Direction of the segment a-b: U when it is closer to the x axis than 45 degrees (space.rs:250-267). 0 = U, 1 = V.
inline int segment_dir(const Point& a, const Point& b) {
    if (CoordKey(a.x).cmp(CoordKey(b.x)) == 0 && CoordKey(a.y).cmp(CoordKey(b.y)) == 0)
        throw MeshError(MeshError::EdgeOnEqualPoints, 0, "Cannot compute the orientation between two Points at the same location");
    const double dx = std::fabs(b.x - a.x), dy = std::fabs(b.y - a.y);
    return std::atan(dy / dx) < 0.78539816339744830961566084581988 ? 0 : 1;
}
inline double segment_len(const Point& a, const Point& b) {  // space.rs:269-274
    const double dx = std::fabs(b.x - a.x), dy = std::fabs(b.y - a.y);
    return std::sqrt(dx * dx + dy * dy);
}

struct Materials { double eps_re = 1, eps_im = 0, mu_re = 1, mu_im = 0; };   // element.rs:82-105
struct Element { uint32_t id; Point points[4]; Materials materials; };       // element.rs:15-19

struct Node { uint32_t id; Point coords; bool boundary; };

struct Edge {   // edge.rs:65-76
    uint32_t id;
    uint32_t nodes[2];
    bool boundary;
    uint8_t dir;          // 0 U, 1 V
    double length;
    int32_t parent = -1;
    int32_t child[2] = {-1, -1};
    int32_t child_node = -1;
    // Registry of adjacent Elems per side.  The reference keeps a BTreeMap<[u8;2], elem> per side and only ever reads the
    // LAST entry (edge.rs:216-222) -> keeping the maximum rank key is equivalent.
    int32_t top_rank[2] = {-1, -1};
    int32_t top_elem[2] = {-1, -1};
    int32_t active[2] = {-1, -1};   // active_elem_pair (edge.rs:193)
    bool has_children() const { return child[0] >= 0; }
    bool has_active_pair() const { return active[0] >= 0; }
};

struct Elem {   // elem.rs:101-110
    uint32_t id;
    uint32_t nodes[4];
    uint32_t edges[4];
    uint32_t element;
    int32_t parent = -1;
    uint8_t loc = BASE;        // HRefLoc inside the parent
    uint8_t h_u = 0, h_v = 0;  // HLevels
    Orders poly_orders;
    int32_t first_child = -1;  // children have consecutive ids (h_refinement.rs:119-134)
    uint8_t n_children = 0;
    double range[4] = {-1.0, 1.0, -1.0, 1.0};   // parametric_range() relative to the Element (elem.rs:191-197), cached
    bool has_children() const { return n_children > 0; }
};

class Mesh {
  public:
    std::vector<Element> elements;
    std::vector<Elem> elems;
    std::vector<Node> nodes;
    std::vector<Edge> edges;

    static Mesh blank() { return Mesh(); }   // mesh.rs:91-98

    // Single [-1,1]^2 cell, unit materials, orders (1,1) (mesh.rs:59-88). The four boundary edges are NOT registered with
    // the Elem, exactly as in the reference.
    static Mesh unit() {
        Mesh m;
        const Point p[4] = {{-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
        Element el; el.id = 0; for (int k = 0; k < 4; k++) el.points[k] = p[k];
        m.elements.push_back(el);
        for (uint32_t k = 0; k < 4; k++) m.nodes.push_back({k, p[k], true});
        const uint32_t en[4][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}};
        for (uint32_t k = 0; k < 4; k++) m.edges.push_back(m.make_edge(k, en[k][0], en[k][1], true));
        Elem e; e.id = 0; e.element = 0;
        for (uint32_t k = 0; k < 4; k++) { e.nodes[k] = k; e.edges[k] = k; }
        m.elems.push_back(e);
        return m;
    }

    // From parsed mesh-file content (mesh.rs:142-327): materials[4*e..], node_ids[4*e..], xy[2*n..].
    static Mesh from_arrays(size_t n_elements, const double* materials, const int64_t* node_ids, size_t n_nodes, const double* xy) {
        Mesh m;
        std::vector<int> uses(n_nodes, 0);
        for (size_t e = 0; e < n_elements; e++)
            for (int k = 0; k < 4; k++) {
                int64_t n = node_ids[4 * e + k];
                if (n < 0 || (size_t)n >= n_nodes) throw MeshError(MeshError::BadMeshFile, e, "node_ids must be smaller than the total number of nodes!");
                for (int q = 0; q < k; q++)
                    if (node_ids[4 * e + q] == n) throw MeshError(MeshError::BadMeshFile, e, "Element's node_ids should have 4 unique values!");
                uses[n]++;
            }
        for (size_t a = 0; a < n_nodes; a++)
            for (size_t b = a + 1; b < n_nodes; b++)
                if (CoordKey(xy[2 * a]).cmp(CoordKey(xy[2 * b])) == 0 && CoordKey(xy[2 * a + 1]).cmp(CoordKey(xy[2 * b + 1])) == 0)
                    throw MeshError(MeshError::BadMeshFile, a, "All Nodes must be at unique locations!");
        for (size_t n = 0; n < n_nodes; n++) {
            if (uses[n] > 4) throw MeshError(MeshError::BadMeshFile, n, "Nodes can only be shared by a maximum of 4 Elements");
            m.nodes.push_back({(uint32_t)n, Point{xy[2 * n], xy[2 * n + 1]}, uses[n] < 4});
        }
        for (size_t e = 0; e < n_elements; e++) {
            Element el; el.id = (uint32_t)e;
            for (int k = 0; k < 4; k++) el.points[k] = m.nodes[node_ids[4 * e + k]].coords;
            el.materials = {materials[4 * e], materials[4 * e + 1], materials[4 * e + 2], materials[4 * e + 3]};
            m.elements.push_back(el);
        }
        // Edge ids = rank of the (node, node) pair in lexicographic order (BTreeMap keys, mesh.rs:202-258).
        // Element-local sides: bottom (nodes 0-1) and left (0-2) see the Element on their top/right side (1); top (2-3) and
        // right (1-3) see it on their bottom/left side (0)  (mesh.rs:1680-1681).
        struct Side { int64_t adj[2] = {-1, -1}; };
        std::map<std::pair<uint32_t, uint32_t>, Side> by_nodes;
        const int side_nodes[4][3] = {{0, 1, 1}, {2, 3, 0}, {0, 2, 1}, {1, 3, 0}};
        for (size_t e = 0; e < n_elements; e++)
            for (auto& s : side_nodes) {
                auto key = std::make_pair((uint32_t)node_ids[4 * e + s[0]], (uint32_t)node_ids[4 * e + s[1]]);
                Side& sd = by_nodes[key];
                if (sd.adj[s[2]] >= 0) throw MeshError(MeshError::BadMeshFile, e, "Edge side has already been set");
                sd.adj[s[2]] = (int64_t)e;
            }
        std::vector<std::array<int64_t, 4>> slots(n_elements, std::array<int64_t, 4>{-1, -1, -1, -1});
        uint32_t eid = 0;
        for (auto& kv : by_nodes) {
            const bool boundary = (kv.second.adj[0] < 0) != (kv.second.adj[1] < 0);
            m.edges.push_back(m.make_edge(eid, kv.first.first, kv.first.second, boundary));
            for (int side = 0; side < 2; side++) {
                if (kv.second.adj[side] < 0) continue;
                // slot of this edge inside the adjacent Elem (mesh.rs:269-275)
                const int slot = m.edges[eid].dir == 0 ? (side == 0 ? 1 : 0) : (side == 0 ? 3 : 2);
                auto& sl = slots[kv.second.adj[side]][slot];
                if (sl >= 0) throw MeshError(MeshError::BadMeshFile, eid, "Elem edge slot has already been set");
                sl = eid;
            }
            eid++;
        }
        for (size_t e = 0; e < n_elements; e++) {
            Elem el; el.id = (uint32_t)e; el.element = (uint32_t)e;
            for (int k = 0; k < 4; k++) {
                if (slots[e][k] < 0) throw MeshError(MeshError::BadMeshFile, e, "Elem is missing an Edge");
                el.nodes[k] = (uint32_t)node_ids[4 * e + k];
                el.edges[k] = (uint32_t)slots[e][k];
            }
            m.elems.push_back(el);
            for (int k = 0; k < 4; k++) m.register_elem_on_edge(el.edges[k], m.elems.back());
        }
        m.set_edge_activation();
        return m;
    }

    static Mesh from_file(const std::string& path);   // json reader in mesh_json.hpp

    // ---- queries ------------------------------------------------------------------------------------------------------
    std::array<Point, 4> elem_points(size_t id) const {   // mesh.rs:376-384
        check_elem(id);
        return {nodes[elems[id].nodes[0]].coords, nodes[elems[id].nodes[1]].coords, nodes[elems[id].nodes[2]].coords, nodes[elems[id].nodes[3]].coords};
    }
    std::vector<size_t> descendant_elems(size_t id, bool include_start) const {   // mesh.rs:470-493 (pre-order DFS)
        check_elem(id);
        std::vector<size_t> out, stack{id};
        while (!stack.empty()) {
            size_t e = stack.back(); stack.pop_back();
            if (e != id || include_start) out.push_back(e);
            for (int k = elems[e].n_children - 1; k >= 0; k--) stack.push_back(elems[e].first_child + k);
        }
        return out;
    }
    std::vector<size_t> ancestor_elems(size_t id, bool include_start) const {     // mesh.rs:520-541
        check_elem(id);
        std::vector<size_t> out;
        if (include_start) out.push_back(id);
        for (int32_t p = elems[id].parent; p >= 0; p = elems[p].parent) out.push_back(p);
        return out;
    }
    std::array<uint8_t, 2> max_expansion_orders() const {   // mesh.rs:626-630
        std::array<uint8_t, 2> r{0, 0};
        for (auto& e : elems) { r[0] = std::max(r[0], e.poly_orders.ni); r[1] = std::max(r[1], e.poly_orders.nj); }
        return r;
    }
    bool elem_is_h_refineable(size_t id) const {   // mesh.rs:641-653
        check_elem(id);
        if (elems[id].has_children()) return false;
        for (uint32_t e : elems[id].edges) if (!(edges[e].length > MIN_EDGE_LENGTH)) return false;
        return true;
    }
    std::array<std::array<int8_t, 2>, 2> elem_p_refinement_window(size_t id) const {   // mesh.rs:666-687
        check_elem(id);
        const Orders& o = elems[id].poly_orders;
        return {{{(int8_t)-(o.ni - 1), (int8_t)(MAX_POLYNOMIAL_ORDER - o.ni)}, {(int8_t)-(o.nj - 1), (int8_t)(MAX_POLYNOMIAL_ORDER - o.nj)}}};
    }
    // Range of `id` inside its ancestor `from_ancestor` (elem.rs:170-188).
    std::array<double, 4> relative_parametric_range(size_t id, size_t from_ancestor) const {
        std::vector<uint8_t> locs;
        int32_t cur = (int32_t)id;
        while (cur >= 0 && (size_t)cur != from_ancestor) { locs.push_back(elems[cur].loc); cur = elems[cur].parent; }
        if (cur < 0) throw MeshError(MeshError::Internal, id, "not a descendant of the given ancestor");
        std::array<double, 4> r{-1.0, 1.0, -1.0, 1.0};
        for (auto it = locs.rbegin(); it != locs.rend(); ++it) { double o[4]; sub_range(*it, r.data(), o); std::copy(o, o + 4, r.begin()); }
        return r;
    }

    // ---- h-refinement (mesh.rs:713-914) ----------------------------------------------------------------------------------
    void global_h_refinement(HRef r) {
        std::vector<std::pair<size_t, HRef>> req;
        for (auto& e : elems) if (elem_is_h_refineable(e.id)) req.push_back({e.id, r});
        execute_h_refinements(req);
    }
    void h_refine_elems(const std::vector<size_t>& ids, HRef r) {
        require_unique(ids);
        std::vector<std::pair<size_t, HRef>> req;
        for (size_t id : ids) req.push_back({id, r});
        execute_h_refinements(req);
    }
    void h_refine_with_filter(const std::function<std::optional<HRef>(const Elem&)>& filt) {
        std::vector<std::pair<size_t, HRef>> req;
        for (auto& e : elems)
            if (elem_is_h_refineable(e.id)) if (auto r = filt(e)) req.push_back({e.id, *r});
        execute_h_refinements(req);
    }
    // Validates everything first ("If any errors are encountered, none of the refinements are executed", mesh.rs:818).
    void execute_h_refinements(const std::vector<std::pair<size_t, HRef>>& requests) {
        std::map<size_t, HRef> merged;   // ascending Elem id
        for (auto& rq : requests) {
            if (rq.first >= elems.size()) throw MeshError(MeshError::ElemDoesNotExist, rq.first, "Elem does not exist; cannot apply h-Refinement!");
            if (!elem_is_h_refineable(rq.first)) throw MeshError(MeshError::ElemNotRefineable, rq.first, "Elem cannot be h-refined; it is either too small or has already been refined!");
            auto it = merged.find(rq.first);
            if (it == merged.end()) merged.emplace(rq.first, rq.second); else it->second.merge(rq.second);
        }
        std::vector<std::pair<size_t, HRef>> follow_ups;
        for (auto& kv : merged) {
            const uint32_t first = split_elem((uint32_t)kv.first, kv.second.kind);
            if (kv.second.kind != HRef::T && kv.second.ext >= 0)
                follow_ups.push_back({first + kv.second.ext, kv.second.kind == HRef::U ? HRef::v() : HRef::u()});
        }
        if (!follow_ups.empty()) execute_h_refinements(follow_ups);
        set_edge_activation();
    }

    // ---- p-refinement (mesh.rs:1265-1665) -----------------------------------------------------------------------------------
    void global_p_refinement(PRef r) {
        std::vector<std::pair<size_t, PRef>> req;
        for (auto& e : elems) req.push_back({e.id, clamp_pref(r, e.id)});
        execute_p_refinements(req);
    }
    void p_refine_elems(const std::vector<size_t>& ids, PRef r) {
        require_unique(ids);
        std::vector<std::pair<size_t, PRef>> req;
        for (size_t id : ids) req.push_back({id, r});
        execute_p_refinements(req);
    }
    void p_refine_with_filter(const std::function<std::optional<PRef>(const Elem&)>& filt) {
        std::vector<std::pair<size_t, PRef>> req;
        for (auto& e : elems) if (auto r = filt(e)) req.push_back({e.id, clamp_pref(*r, e.id)});
        execute_p_refinements(req);
    }
    void execute_p_refinements(const std::vector<std::pair<size_t, PRef>>& requests) {
        std::map<size_t, std::array<int, 2>> merged;
        for (auto& rq : requests) {
            if (rq.first >= elems.size()) throw MeshError(MeshError::ElemDoesNotExist, rq.first, "Elem does not exist; cannot apply p-Refinement!");
            auto& d = merged[rq.first];   // zero-initialised on first touch
            d[0] += rq.second.di; d[1] += rq.second.dj;
        }
        for (auto& kv : merged) {
            auto w = elem_p_refinement_window(kv.first);
            if (kv.second[0] < w[0][0] || kv.second[0] > w[0][1] || kv.second[1] < w[1][0] || kv.second[1] > w[1][1])
                throw MeshError(MeshError::RefinementOutOfBounds, kv.first, "p-Refinement out of bounds");
        }
        for (auto& kv : merged) {
            elems[kv.first].poly_orders.ni = (uint8_t)(elems[kv.first].poly_orders.ni + kv.second[0]);
            elems[kv.first].poly_orders.nj = (uint8_t)(elems[kv.first].poly_orders.nj + kv.second[1]);
        }
    }
    void set_global_expansion_orders(Orders o) {
        std::vector<std::pair<size_t, Orders>> req;
        for (auto& e : elems) req.push_back({e.id, o});
        set_expansion_orders(req);
    }
    void set_expansion_on_elems(const std::vector<size_t>& ids, Orders o) {
        std::vector<std::pair<size_t, Orders>> req;
        for (size_t id : ids) req.push_back({id, o});
        set_expansion_orders(req);
    }
    void set_expansions_with_filter(const std::function<std::optional<Orders>(const Elem&)>& filt) {
        std::vector<std::pair<size_t, Orders>> req;
        for (auto& e : elems) if (auto o = filt(e)) req.push_back({e.id, *o});
        set_expansion_orders(req);
    }
    void set_expansion_orders(const std::vector<std::pair<size_t, Orders>>& requests) {   // mesh.rs:1646-1665
        std::map<size_t, Orders> merged;
        for (auto& rq : requests) {
            if (rq.first >= elems.size()) throw MeshError(MeshError::ElemDoesNotExist, rq.first, "Elem does not exist");
            if (!merged.emplace(rq.first, rq.second).second) throw MeshError(MeshError::DuplicateElemIds, rq.first, "Duplicate element ids");
        }
        for (auto& kv : merged) {   // PolyOrders::set p_refinement.rs:30-42
            if (kv.second.ni > MAX_POLYNOMIAL_ORDER || kv.second.nj > MAX_POLYNOMIAL_ORDER) throw MeshError(MeshError::ExceededMaxExpansion, kv.first, "Exceeded maximum expansion order");
            if (kv.second.ni < 1 || kv.second.nj < 1) throw MeshError(MeshError::NegExpansion, kv.first, "Expansion orders must be at least 1");
        }
        for (auto& kv : merged) elems[kv.first].poly_orders = kv.second;
    }

    // ---- edge activation (mesh.rs:1180-1215) ------------------------------------------------------------------------------------
    // An edge carries edge-type DoFs for the highest-ranked Elem on each side; a parent edge whose two halves are both
    // supported hands over to them.
    void set_edge_activation() {
        for (auto& e : edges) e.active[0] = e.active[1] = -1;
        for (size_t k = 0; k < edges.size(); k++)
            if (edges[k].parent < 0 && !edges[k].boundary)
                if (!activate_tree((uint32_t)k)) throw MeshError(MeshError::Internal, k, "Unable to find active Edge pair; Something must be wrong with the mesh!");
    }

  private:
    void check_elem(size_t id) const { if (id >= elems.size()) throw MeshError(MeshError::ElemDoesNotExist, id, "Attempt to access non-existent elem"); }
    static void require_unique(const std::vector<size_t>& ids) {
        std::vector<size_t> s(ids); std::sort(s.begin(), s.end());
        if (std::adjacent_find(s.begin(), s.end()) != s.end()) throw MeshError(MeshError::DuplicateElemIds, 0, "Duplicate element ids in refinement");
    }
    PRef clamp_pref(PRef r, size_t id) const {   // PRef::constrained_to
        auto w = elem_p_refinement_window(id);
        return {(int8_t)std::min<int>(std::max<int>(r.di, w[0][0]), w[0][1]), (int8_t)std::min<int>(std::max<int>(r.dj, w[1][0]), w[1][1])};
    }
    Edge make_edge(uint32_t id, uint32_t n0, uint32_t n1, bool boundary) const {   // edge.rs:80-95
        Edge e; e.id = id; e.nodes[0] = n0; e.nodes[1] = n1; e.boundary = boundary;
        e.dir = (uint8_t)segment_dir(nodes[n0].coords, nodes[n1].coords);
        e.length = segment_len(nodes[n0].coords, nodes[n1].coords);
        return e;
    }
    // edge.rs:97-123: side 1 (top / right) when the edge is the Elem's slot 0 or 2; rank = HLevels ordered with the across-edge
    // level first (h_refinement.rs:31-36).
    void register_elem_on_edge(uint32_t edge_id, const Elem& el) {
        Edge& ed = edges[edge_id];
        int slot = -1;
        for (int k = 0; k < 4; k++) if (el.edges[k] == edge_id) { slot = k; break; }
        if (slot < 0) throw MeshError(MeshError::Internal, el.id, "Elem is not connected to Edge; cannot reciprocate connection!");
        const int side = (slot == 0 || slot == 2) ? 1 : 0;
        const int rank = ed.dir == 0 ? (el.h_v << 8 | el.h_u) : (el.h_u << 8 | el.h_v);
        if (rank > ed.top_rank[side]) { ed.top_rank[side] = rank; ed.top_elem[side] = (int32_t)el.id; }
        else if (rank == ed.top_rank[side] && ed.top_elem[side] != (int32_t)el.id)
            throw MeshError(MeshError::Internal, el.id, "Edge is already connected to another Elem at this rank");
    }
    bool activate_tree(uint32_t k) {
        Edge& e = edges[k];
        if (e.top_elem[0] < 0 || e.top_elem[1] < 0) { e.active[0] = e.active[1] = -1; return false; }
        e.active[0] = e.top_elem[0]; e.active[1] = e.top_elem[1];
        if (e.has_children()) {
            const uint32_t c0 = (uint32_t)e.child[0], c1 = (uint32_t)e.child[1];
            const bool a = activate_tree(c0), b = activate_tree(c1);
            if (a != b) throw MeshError(MeshError::Internal, k, "Children of Edge do not have consistent support for Basis Functions");
            if (a) { edges[k].active[0] = edges[k].active[1] = -1; }
        }
        return true;
    }
    // Bisect an edge unless it already is (mesh.rs:1095-1126, edge.rs:126-170): two new edge ids first, then one node id.
    struct Bisection { uint32_t half[2]; uint32_t mid; };
    Bisection bisect_edge(uint32_t k) {
        if (edges[k].has_children()) return {{(uint32_t)edges[k].child[0], (uint32_t)edges[k].child[1]}, (uint32_t)edges[k].child_node};
        const double half_len = edges[k].length / 2.0;
        if (half_len < MIN_EDGE_LENGTH) throw MeshError(MeshError::MinEdgeLength, k, "h-refinement will result in Edge length below minimum");
        const uint32_t e0 = (uint32_t)edges.size(), e1 = e0 + 1, mid = (uint32_t)nodes.size();
        const Edge par = edges[k];
        nodes.push_back({mid, Point::mid(nodes[par.nodes[0]].coords, nodes[par.nodes[1]].coords), par.boundary});
        Edge h; h.boundary = par.boundary; h.dir = par.dir; h.length = half_len; h.parent = (int32_t)k;
        h.id = e0; h.nodes[0] = par.nodes[0]; h.nodes[1] = mid; edges.push_back(h);
        h.id = e1; h.nodes[0] = mid; h.nodes[1] = par.nodes[1]; edges.push_back(h);
        edges[k].child[0] = (int32_t)e0; edges[k].child[1] = (int32_t)e1; edges[k].child_node = (int32_t)mid;
        return {{e0, e1}, mid};
    }
    // New interior (non-boundary, parentless) edge; its nodes are ordered along its direction (mesh.rs:1128-1156).
    uint32_t interior_edge(uint32_t na, uint32_t nb, size_t parent_elem) {
        const Point& a = nodes[na].coords; const Point& b = nodes[nb].coords;
        const int c = segment_dir(a, b) == 0 ? CoordKey(a.x).cmp(CoordKey(b.x)) : CoordKey(a.y).cmp(CoordKey(b.y));
        if (c == 0) throw MeshError(MeshError::EdgeOnEqualPoints, parent_elem, "Attempt to generate a child-Edge between two identical points");
        const uint32_t id = (uint32_t)edges.size();
        edges.push_back(c < 0 ? make_edge(id, na, nb, false) : make_edge(id, nb, na, false));
        return id;
    }
    // Creates the 4 (T) or 2 (U/V) children of `pid`; returns the id of the first child.  Id allocation order follows the
    // reference exactly: children ids, [centre node], then per parent edge: bisection ids, interior edge id.
    uint32_t split_elem(uint32_t pid, HRef::Kind kind) {
        if (elems[pid].has_children()) throw MeshError(MeshError::ElemHasChildren, pid, "Elem already has children; Cannot h-refine!");
        const int nc = kind == HRef::T ? 4 : 2;
        const uint32_t first = (uint32_t)elems.size();
        const Elem par = elems[pid];
        std::vector<Elem> ch(nc);
        static const uint8_t locs[3][4] = {{SW, SE, NW, NE}, {W, E, 0, 0}, {S, N, 0, 0}};
        for (int k = 0; k < nc; k++) {
            Elem& c = ch[k];
            c.id = first + k; c.element = par.element; c.parent = (int32_t)pid; c.loc = locs[kind][k];
            c.h_u = par.h_u + (kind != HRef::V); c.h_v = par.h_v + (kind != HRef::U);   // h_refinement.rs:22-28
            c.poly_orders = par.poly_orders;                                             // elem.rs:146
            sub_range(c.loc, par.range, c.range);
            for (int q = 0; q < 4; q++) { c.nodes[q] = UINT32_MAX; c.edges[q] = UINT32_MAX; }
        }
        if (kind == HRef::T) {   // mesh.rs:916-979
            const uint32_t centre = (uint32_t)nodes.size();
            nodes.push_back({centre, Point::mid(nodes[par.nodes[0]].coords, nodes[par.nodes[3]].coords), false});
            for (int k = 0; k < 4; k++) { ch[k].nodes[3 - k] = centre; ch[k].nodes[k] = par.nodes[k]; }
            // per parent edge slot: the two children along it (in edge direction), the child node slot that receives the
            // midpoint, and the child edge slot that receives the new interior edge
            static const uint8_t lay[4][6] = {{0, 1, 1, 0, 3, 2}, {2, 3, 3, 2, 3, 2}, {0, 2, 2, 0, 1, 0}, {1, 3, 3, 1, 1, 0}};
            for (int s = 0; s < 4; s++) {
                const Bisection b = bisect_edge(par.edges[s]);
                Elem& ca = ch[lay[s][0]]; Elem& cb = ch[lay[s][1]];
                ca.edges[s] = b.half[0]; cb.edges[s] = b.half[1];
                ca.nodes[lay[s][2]] = b.mid; cb.nodes[lay[s][3]] = b.mid;
                const uint32_t ie = interior_edge(b.mid, centre, pid);
                ca.edges[lay[s][4]] = ie; cb.edges[lay[s][5]] = ie;
            }
        } else {   // U: bisect slots 0,1 (mesh.rs:981-1036); V: bisect slots 2,3 (mesh.rs:1038-1093)
            const int s0 = kind == HRef::U ? 0 : 2;
            uint32_t mids[2];
            for (int q = 0; q < 2; q++) {
                const int s = s0 + q;
                const Bisection b = bisect_edge(par.edges[s]);
                mids[q] = b.mid;
                ch[0].edges[s] = b.half[0]; ch[1].edges[s] = b.half[1];
                // node slots on that parent edge: U: slot0 -> nodes (0,1), slot1 -> nodes (2,3); V: slot2 -> (0,2), slot3 -> (1,3)
                const int n_lo = kind == HRef::U ? 2 * q : q, n_hi = kind == HRef::U ? 2 * q + 1 : q + 2;
                ch[0].nodes[n_lo] = par.nodes[n_lo]; ch[0].nodes[n_hi] = b.mid;
                ch[1].nodes[n_lo] = b.mid;           ch[1].nodes[n_hi] = par.nodes[n_hi];
            }
            const uint32_t ie = interior_edge(mids[0], mids[1], pid);
            if (kind == HRef::U) { ch[0].edges[3] = ie; ch[1].edges[2] = ie; ch[0].edges[2] = par.edges[2]; ch[1].edges[3] = par.edges[3]; }
            else                 { ch[0].edges[1] = ie; ch[1].edges[0] = ie; ch[0].edges[0] = par.edges[0]; ch[1].edges[1] = par.edges[1]; }
        }
        for (auto& c : ch) for (int q = 0; q < 4; q++)
            if (c.nodes[q] == UINT32_MAX || c.edges[q] == UINT32_MAX) throw MeshError(MeshError::Internal, c.id, "child Elem was not fully initialised");
        elems[pid].first_child = (int32_t)first; elems[pid].n_children = (uint8_t)nc;
        for (auto& c : ch) { elems.push_back(c); }
        for (int k = 0; k < nc; k++) for (int q = 0; q < 4; q++) register_elem_on_edge(elems[first + k].edges[q], elems[first + k]);   // mesh.rs:1167-1171
        return first;
    }
};

}  // namespace fem2d
