"""Host logic of the product (no GPU): the Mesh/Domain mirror, the flattened view, the symbolic phase and the C-ABI surface,
checked against the CPU oracle and the reference's doctests."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import fem_2d_b200 as F
import oracle as O
import recipes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    """The shared library must export exactly what include/*.h declares (no compute calls here)."""
    declared = set()
    for hdr in ("fem2d.h", "fem2d_host.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        declared |= set(re.findall(r"\b(fem2dh?_[a-z0-9_]+)\s*\(", txt))
    declared -= {"fem2d_domain_view", "fem2d_plan", "fem2dh_mesh", "fem2dh_domain"}
    assert declared == set(F.ABI_SYMBOLS) | set(F.HOST_ABI_SYMBOLS), declared ^ (set(F.ABI_SYMBOLS) | set(F.HOST_ABI_SYMBOLS))
    lib = C.CDLL(F.LIB_PATH)
    for s in sorted(declared):
        assert hasattr(lib, s), s
    assert F.version().startswith("fem2d-b200")
    assert isinstance(F.device_count(), int)


@pytest.mark.parametrize("name", sorted(recipes.RECIPES))
def test_mesh_and_domain_match_oracle(name):
    mo, mf = recipes.build_pair(name)
    assert (mo.num_elems, mo.num_edges, mo.num_nodes) == (mf.num_elems, mf.num_edges, mf.num_nodes)
    for e in range(mo.num_elems):
        a, b = mo.elem(e), mf.elem(e)
        assert (a.nodes, a.edges, a.parent, a.children, a.ni, a.nj, a.h_u, a.h_v, a.element) == (b.nodes, b.edges, b.parent, b.children, b.ni, b.nj, b.h_u, b.h_v, b.element)
        assert mo.parametric_range(e) == mf.parametric_range(e)
    for e in range(mo.num_edges):
        assert mo.edge(e) == mf.edge(e)
    for n in range(mo.num_nodes):
        assert mo.node(n) == mf.node(n)
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    assert do.num_dofs == df.num_dofs
    for e in range(mo.num_elems):
        for x, y in zip(do.local_basis_specs(e), df.local_basis_specs(e)):
            assert np.array_equal(x, y)
    # flattened view
    v = df.view()
    assert v.n_elems == mf.num_elems and v.n_dofs == df.num_dofs and [v.i_max, v.j_max] == mf.max_expansion_orders()
    for e in range(mf.num_elems):
        i, j, d, dof = df.local_basis_specs(e)
        s = slice(int(v.bs_off[e]), int(v.bs_off[e + 1]))
        assert np.array_equal(v.bs_i[s], i) and np.array_equal(v.bs_j[s], j) and np.array_equal(v.bs_dir[s], d) and np.array_equal(v.bs_dof[s], dof)
        assert v.elem_parent[e] == mf.elem(e).parent
    # relative ranges of every (ancestor, descendant) pair
    for e in range(min(mf.num_elems, 40)):
        for a in mf.ancestor_elems(e, False):
            assert mo.parametric_range(e, a) == mf.parametric_range(e, a)


@pytest.mark.parametrize("name", sorted(recipes.RECIPES))
def test_host_symbolic_pattern_equals_oracle_key_set(name):
    mo, mf = recipes.build_pair(name)
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    ref = O.galerkin_sample_gep_hcurl(do, [4, 4])
    for dedupe in (True, False):
        plan = F.Plan(df.view(), device=-1, dedupe=dedupe)
        rows, cols = plan.pattern()
        assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
        assert np.all(rows <= cols)
        keys = rows.astype(np.uint64) << np.uint64(32) | cols.astype(np.uint64)
        assert np.all(np.diff(keys.astype(np.int64)) > 0)          # BTreeMap<[u32;2]> order, unique
        assert plan.info["max_contrib"] <= 2                        # <= 2 contributions per key => order-independent sums
        assert plan.info["n_pairs"] - plan.info["nnz_upper"] == plan.info["n_extra"]
    # reference pair count: sum_e n_e(n_e+1)/2 + sum_{e, d in desc(e)} n_e n_d  (galerkin.rs:91-178)
    n_pairs = 0
    for e in range(mf.num_elems):
        ne = len(df.local_basis_specs(e)[0])
        n_pairs += ne * (ne + 1) // 2 + sum(ne * len(s[1][0]) for s in df.descendant_basis_specs(e))
    assert plan.info["n_pairs"] == n_pairs


@pytest.mark.parametrize("name", sorted(recipes.RECIPES))
def test_work_items_cover_every_pair_once(name):
    """The exact integrator's decomposition (plan_types.h): every pair the pattern reads lies in exactly one micro-tile
    (4 x 2 / 1 x 2 same-direction, 4 x 4 / 1 x 4 cross-direction), every tile in exactly one work item, same-direction tiles
    first, staged function ranges cover what the tiles read -- on every recipe, both dedupe settings (checked on the host)."""
    _, mf = recipes.build_pair(name)
    df = F.Domain.from_mesh(mf)
    for dedupe in (True, False):
        plan = F.Plan(df.view(), device=-1, dedupe=dedupe)
        chk = plan.check_work_items()
        assert chk["violations"] == 0, chk
        assert 0 < chk["same_tiles"] < chk["tiles"] <= chk["slots"] < chk["tiles"] + 32 * plan.info["n_work_items"]


def test_work_items_of_the_baseline_configs():
    """BASELINE.json configs[1] (order 8, 144 functions per Elem) and configs[3] (anisotropic, p in [2, 10], local-desc blocks of
    unequal lists): the decomposition is consistent in both tile shapes."""
    import bench
    for wl, tile_p in (("cfg2", 1), ("cfg4", 4)):
        df = bench.build_product_domain(wl)
        for dedupe in (True, False):
            plan = F.Plan(df.view(), device=-1, dedupe=dedupe)
            assert not dedupe or plan.info["tile_p"] == tile_p
            chk = plan.check_work_items()
            assert chk["violations"] == 0, (wl, dedupe, chk)


def test_work_items_of_the_throughput_shape():
    """A plan large enough for the 4-row tile shape (cfg 3 at 3 levels, dedupe off): closed-form tile counts of an interior leaf
    (42 U + 42 V functions: 121 + 121 same-direction tiles of 4 x 2 in the two triangles, 11 x 11 cross-direction tiles of 4 x 4)."""
    df = F.Domain.from_mesh(recipes.mesh_cfg3(recipes.api("product"), levels=3, order=6))
    plan = F.Plan(df.view(), device=-1, dedupe=False)
    assert plan.info["tile_p"] == 4
    chk = plan.check_work_items()
    assert chk["violations"] == 0, chk
    n_int = 14 * 14                                            # leaves with four interior edges on the 16 x 16 leaf grid
    assert chk["tiles"] > n_int * 363 and chk["same_tiles"] > n_int * 242


def test_baseline_config_sizes():
    """BASELINE.md section 3 table (cfg 1, 5 fully; cfg 3 closed form at a small level)."""
    for name, (elems, dofs, pairs, nnz) in {"readme": (28, 600, 12808 + 1152, 13596), "slepc": (27, 624, 12868 + 3328, 15512)}.items():
        mf = recipes.RECIPES[name](recipes.api("product"))
        df = F.Domain.from_mesh(mf)
        plan = F.Plan(df.view(), device=-1)
        assert (mf.num_elems, df.num_dofs, plan.info["n_pairs"], plan.nnz) == (elems, dofs, pairs, nnz)
    L = 3
    mf = recipes.mesh_cfg3(recipes.api("product"), levels=L, order=6)
    df = F.Domain.from_mesh(mf)
    nx = ny = 2 * 2 ** L
    n_int = nx * (ny - 1) + (nx - 1) * ny
    assert df.num_dofs == 60 * nx * ny + 6 * n_int                    # SURVEY.md section 8 closed form
    plan = F.Plan(df.view(), device=-1)
    assert plan.nnz == plan.info["n_pairs"] - 21 * n_int


def test_row_blocks_are_row_aligned_and_balanced():
    df = F.Domain.from_mesh(recipes.mesh_cfg3(recipes.api("product"), levels=2, order=4))
    plan = F.Plan(df.view(), device=-1)
    rows, _ = plan.pattern()
    for world in (1, 2, 3, 8):
        b = plan.row_blocks(world)
        assert b[0] == 0 and b[-1] == plan.nnz and np.all(np.diff(b.astype(np.int64)) >= 0)
        for k in b[1:-1]:
            assert rows[int(k)] != rows[int(k) - 1]
        assert np.max(np.diff(b.astype(np.int64))) < plan.nnz / world + 500


def test_row_offsets_are_the_csr_form_of_the_pattern():
    df = F.Domain.from_mesh(recipes.build_pair("slepc")[1])
    plan = F.Plan(df.view(), device=-1)
    rows, _ = plan.pattern()
    rp = plan.row_offsets()
    assert len(rp) == df.num_dofs + 1 and rp[0] == 0 and rp[-1] == plan.nnz
    assert np.array_equal(np.repeat(np.arange(df.num_dofs, dtype=np.uint32), np.diff(rp).astype(np.int64)), rows)


def test_error_behaviour_mirrors_reference():
    m = F.Mesh.unit()
    m.h_refine_elems([0], F.HRef.T)
    m.h_refine_elems([2, 3, 4], F.HRef.v())
    assert m.num_elems == 11                                           # mesh.rs:740-751
    for bad, kind in (([15], "ElemDoesNotExist"), ([0], "ElemNotRefineable"), ([1, 1], "DuplicateElemIds")):
        with pytest.raises(F.MeshError) as e:
            m.h_refine_elems(bad, F.HRef.T)
        assert e.value.kind == kind
    assert m.num_elems == 11
    with pytest.raises(F.MeshError):
        F.HRef.u_extened(2)                                            # mesh.rs:2007-2014
    m = F.Mesh.from_file(recipes.MESH_C)
    m.set_global_expansion_orders(F.Orders.new(3, 3))
    with pytest.raises(F.MeshError) as e:
        m.p_refine_elems([0], F.PRef(-3, 1))                           # mesh.rs:2047-2053
    assert e.value.kind == "RefinementOutOfBounds"
    with pytest.raises(F.MeshError):
        F.Mesh.from_file(recipes.MESH_C).p_refine_elems([0], F.PRef(20, 0))
    with pytest.raises(F.MeshError) as e:
        mm = F.Mesh.from_file(recipes.MESH_C)
        for _ in range(18):                                            # mesh.rs:2016-2031 minimum edge length
            mm.h_refine_with_filter(lambda el: F.HRef.T if (not el.has_children() and el.nodes[0] == 0) else None)
    # Galerkin errors are decided before any compute and before the device is touched (galerkin.rs:42-59)
    with pytest.raises(F.GalerkinSamplingError) as e:
        F.galerkin_sample_gep_hcurl(F.Domain.blank(F.ContinuityCondition.HDiv), [8, 8])
    assert e.value.kind == F.GalerkinSamplingError.WrongContinuityCondition
    with pytest.raises(F.GalerkinSamplingError) as e:
        F.galerkin_sample_gep_hcurl(F.Domain.blank(F.ContinuityCondition.HCurl), [8, 8])
    assert e.value.kind == F.GalerkinSamplingError.EmptyDOFSet
    with pytest.raises(F.GalerkinSamplingError) as e:      # Domain::unit has orders (1,1) and only boundary edges: no DoFs; checked before GLQ
        F.galerkin_sample_gep_hcurl(F.Domain.unit(), [8, 3])
    assert e.value.kind == F.GalerkinSamplingError.EmptyDOFSet
    d = F.Domain.from_mesh(recipes.mesh_nalg(recipes.api("product")))
    with pytest.raises(F.GalerkinSamplingError) as e:
        F.galerkin_sample_gep_hcurl(d, [8, 3])
    assert e.value.kind == F.GalerkinSamplingError.InvalidGLQSettings


def test_no_cpu_fallback():
    """Without a CUDA device the numeric entry points must refuse, never compute on the host."""
    d = F.Domain.from_mesh(recipes.mesh_nalg(recipes.api("product")))
    glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
    plan = F.Plan(d.view(), device=-1)
    with pytest.raises(F.BackendError) as e:
        plan.assemble(glq)
    assert e.value.status == F.ERR_NO_DEVICE
    if F.device_count() == 0:
        with pytest.raises(F.BackendError) as e:
            F.galerkin_sample_gep_hcurl(d, [8, 8])
        assert e.value.status == F.ERR_NO_DEVICE


def test_glq_mirror_matches_oracle_and_table():
    for n in (4, 5, 8, 12, 16, 20, 32, 64):
        p, w = F.gauss_quadrature_points(n)
        po, wo = O.gauss_quadrature_points(n)
        assert np.max(np.abs(p - po)) < 5e-15 and np.max(np.abs(w - wo)) < 5e-15
        assert abs(w.sum() - 2.0) < 1e-14 and np.all(np.diff(p) > 0)
    assert [F.default_ngq(k) for k in (3, 4, 6, 8, 10)] == [16, 16, 32, 32, 64]


def test_petsc_aij_writer(tmp_path):
    """sparse_matrix.rs:184-264: header 1211216, dims, nnz, per-row counts, sorted column ids, values; all big-endian."""
    rows = np.array([0, 0, 1, 2], dtype=np.uint32); cols = np.array([0, 2, 1, 2], dtype=np.uint32)
    vals = np.array([1.0, 2.5, -3.0, 4.0])
    sm = F.SparseMatrix(3, rows, cols, vals)
    assert sm.num_entries() == 5
    path = str(tmp_path / "m.dat")
    sm.print_to_petsc_binary_file(path)
    raw = open(path, "rb").read()
    hdr = struct.unpack(">4I", raw[:16])
    assert hdr == (1211216, 3, 3, 5) and raw[:4] == b"\x00\x12\x7b\x50"
    counts = struct.unpack(">3I", raw[16:28]); js = struct.unpack(">5I", raw[28:48]); a = struct.unpack(">5d", raw[48:88])
    assert counts == (2, 1, 2) and js == (0, 2, 1, 0, 2) and a == (1.0, 2.5, -3.0, 2.5, 4.0)
    dense = sm.to_dense()
    assert np.array_equal(dense, dense.T) and dense[2, 0] == 2.5


def test_petsc_aij_writer_matches_oracle_restatement(tmp_path):
    """Host writer of the mirror vs the oracle-side restatement of sparse_matrix.rs:184-264 on a real (small) assembled pattern."""
    import oracle as O
    import recipes
    mo = recipes.RECIPES["nalg"](recipes.api("oracle"))
    do = O.Domain.from_mesh(mo)
    gep = O.galerkin_sample_gep_hcurl(do, [4, 4])
    path = str(tmp_path / "m.dat")
    F.SparseMatrix(do.num_dofs, gep.rows, gep.cols, gep.a).print_to_petsc_binary_file(path)
    assert open(path, "rb").read() == O.petsc_aij_bytes(do.num_dofs, gep.rows, gep.cols, gep.a)


def test_planner_folds_only_exact_power_of_two_scales(tmp_path):
    """HostPlan::ws_fold (the scales the persistent integrator may multiply into the quadrature weights, bit for bit): both the uv / vu
    ratios and max(det) on dyadic Element sides (mesh a, 1.0 x 0.5), the ratios only on a 3 x 3 Element (max(det) = 2.25 / 4^k), nothing on
    mesh c (4.2 x 4.2: the refined sides are rounded, some ratios are an ulp off 1), and nothing when the tuning switch says so.  Host-only plans."""
    import json
    mesh3 = str(tmp_path / "mesh_3x3.json")
    with open(mesh3, "w") as f:
        json.dump({"Elements": [{"materials": [1.0, 0.0, 2.0, 0.0], "node_ids": [0, 1, 2, 3]}], "Nodes": [[0.0, 0.0], [3.0, 0.0], [0.0, 3.0], [3.0, 3.0]]}, f)
    api = recipes.api("product")

    def refined(path, order, levels):
        m = api.Mesh.from_file(path)
        api.set_orders(m, order, order)
        for _ in range(levels):
            m.global_h_refinement(api.href(recipes.T))
        return F.Domain.from_mesh(m)

    cases = [(refined(recipes.MESH_A, 9, 2), 3), (refined(mesh3, 12, 2), 1), (refined(recipes.MESH_C, 12, 2), 0)]
    for dom, want in cases:
        plan = F.Plan(dom.view(), device=-1, dedupe=False)
        assert plan.info["tile_p"] == 4, "the throughput tile shape is what folds"
        assert plan.round_fill()["fold"] == want
        assert plan.check_work_items()["violations"] == 0
    os.environ["FEM2D_K2_WS_FOLD"] = "0"
    try:
        assert F.Plan(cases[0][0].view(), device=-1, dedupe=False).round_fill()["fold"] == 0
    finally:
        os.environ.pop("FEM2D_K2_WS_FOLD", None)
