"""Domain recipes shared by the tests, the bench and smoke(): each builds the same mesh through the oracle API
(`lib="oracle"`) or the product host mirror (`lib="product"`).  Recipes follow SURVEY.md section 8 / BASELINE.json configs."""
from __future__ import annotations

import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MESH_A = os.path.join(GOLDEN, "test_mesh_a.json")
MESH_B = os.path.join(GOLDEN, "test_mesh_b.json")
MESH_C = os.path.join(GOLDEN, "test_mesh_c.json")

M64 = (1 << 64) - 1
SEED = 20261017


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


class _Api:
    """Uniform facade over the two Mesh implementations."""

    def __init__(self, lib: str):
        self.lib = lib
        if lib == "oracle":
            import oracle as O
            self.mod = O
            self.Mesh, self.Domain = O.Mesh, O.Domain
        else:
            import fem_2d_b200 as F
            self.mod = F
            self.Mesh, self.Domain = F.Mesh, F.Domain

    def href(self, kind, ext=-1):
        return self.mod.HRef(kind, ext)

    def set_orders(self, m, ni, nj):
        if self.lib == "oracle":
            m.set_global_expansion_orders(ni, nj)
        else:
            m.set_global_expansion_orders(self.mod.Orders(ni, nj))

    def set_orders_list(self, m, lst):
        if self.lib == "oracle":
            m.set_expansion_orders(lst)
        else:
            m.set_expansion_orders([(e, self.mod.Orders(a, b)) for e, a, b in lst])

    def p_refine(self, m, di, dj):
        if self.lib == "oracle":
            m.global_p_refinement(di, dj)
        else:
            m.global_p_refinement(self.mod.PRef(di, dj))

    def p_refine_elems(self, m, ids, di, dj):
        if self.lib == "oracle":
            m.p_refine_elems(ids, di, dj)
        else:
            m.p_refine_elems(ids, self.mod.PRef(di, dj))


T, U, V = 0, 1, 2


def mesh_nalg(api):          # lib.rs:45-53
    m = api.Mesh.from_file(MESH_A)
    api.p_refine(m, 2, 2)
    return m


def mesh_slepc(api):         # lib.rs:85-88 (BASELINE cfg 5)
    m = api.Mesh.from_file(MESH_B)
    api.p_refine(m, 3, 3)
    m.global_h_refinement(api.href(T))
    m.h_refine_elems([6, 9, 12], api.href(T))
    return m


def mesh_readme(api):        # README.md:45-61 (BASELINE cfg 1)
    m = api.Mesh.from_file(MESH_A)
    api.set_orders(m, 4, 4)
    m.global_h_refinement(api.href(T))
    cn = m.elem(0).nodes[3]
    m.h_refine_with_filter(lambda e: api.href(U) if cn in e.nodes else None)
    return m


def mesh_cfg2(api, levels=3, order=8):   # BASELINE cfg 2: mesh_b, Orders(8,8), 3x global T
    m = api.Mesh.from_file(MESH_B)
    api.set_orders(m, order, order)
    for _ in range(levels):
        m.global_h_refinement(api.href(T))
    return m


def mesh_cfg3(api, levels=6, order=6):   # BASELINE cfg 3: mesh_a, Orders(6,6), 6x global T (~1M DoFs)
    m = api.Mesh.from_file(MESH_A)
    api.set_orders(m, order, order)
    for _ in range(levels):
        m.global_h_refinement(api.href(T))
    return m


def mesh_cfg4(api, t_levels=4, rounds=3, pmin=2, pmax=10, seed=SEED, mesh_file=MESH_C):   # BASELINE cfg 4 (SURVEY.md 8d recipe)
    SEED = seed
    m = api.Mesh.from_file(mesh_file)
    for _ in range(t_levels):
        m.global_h_refinement(api.href(T))
    for r in range(rounds):
        def filt(e, r=r):
            h = splitmix64(SEED ^ (e.id * 1000003 + r))
            if h % 2 != 0:
                return None
            return api.href(U) if (h >> 8) & 1 else api.href(V)
        m.h_refine_with_filter(filt)
    span = pmax - pmin + 1
    api.set_orders_list(m, [(e, pmin + splitmix64(SEED ^ (2 * e + (1 << 32))) % span, pmin + splitmix64(SEED ^ (2 * e + 1 + (1 << 32))) % span)
                            for e in range(m.num_elems)])
    return m


def mesh_hp1m(api, t_levels=6, rounds=4):
    """north_star target workload (BASELINE.json metric: ">= 1M-DoF anisotropically hp-refined H(curl) domain"): the cfg-4 recipe
    at 6 global T-levels and 4 seeded U/V rounds, p in [2, 10] per Elem: 38 509 Elems, 1 380 549 DoFs, 84 930 129 upper-triangular
    entries per matrix, 71 086 pair blocks (local-local and local-desc) in 63 157 classes."""
    return mesh_cfg4(api, t_levels=t_levels, rounds=rounds)


def mesh_edge_order(api):    # mesh.rs:1918-1937 recipe + anisotropic orders: exercises U(Some(1)) extensions and n-irregular edges
    m = api.Mesh.from_file(MESH_B)
    m.global_h_refinement(api.href(T))
    be = [m.edge(i)["boundary"] for i in range(m.num_edges)]
    m.h_refine_with_filter(lambda e: api.href(U) if any(be[x] for x in e.edges) else None)
    m.h_refine_with_filter(lambda e: api.href(T) if 4 in e.nodes else api.href(V))
    m.global_h_refinement(api.href(U, 1))
    api.p_refine(m, 2, 1)
    return m


def mesh_create_domain(api):  # domain.rs:399-412
    m = api.Mesh.from_file(MESH_A)
    api.set_orders(m, 5, 5)
    m.global_h_refinement(api.href(T))
    m.h_refine_elems([4, 5], api.href(T))
    m.h_refine_elems([6, 7], api.href(U))
    m.h_refine_elems([8, 9], api.href(V))
    api.p_refine_elems(m, [10, 11, 12, 13], 2, -1)
    return m


RECIPES = {
    "nalg": mesh_nalg, "slepc": mesh_slepc, "readme": mesh_readme, "edge_order": mesh_edge_order,
    "create_domain": mesh_create_domain,
    "cfg2_small": lambda api: mesh_cfg2(api, levels=1, order=5),
    "cfg3_small": lambda api: mesh_cfg3(api, levels=2, order=4),
    "cfg4_small": lambda api: mesh_cfg4(api, t_levels=1, rounds=3, pmin=2, pmax=5),
}


def build_pair(name: str):
    """(oracle mesh, product mesh) of one recipe."""
    return RECIPES[name](_Api("oracle")), RECIPES[name](_Api("product"))


def api(lib: str) -> _Api:
    return _Api(lib)
