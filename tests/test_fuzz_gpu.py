"""Randomised differential test of the whole CUDA path against the oracle: seeded anisotropic U/V refinements to n-irregularity up to 3
with random per-Elem orders on all three reference meshes (the cfg-4 recipe family of SURVEY.md 8d with varying seeds).  Every case checks
the device pattern (symbolic assembly by rows + sorted shared rows) and the A/B bits, with and without block dedupe."""
import numpy as np
import pytest

import recipes

pytestmark = pytest.mark.gpu

import fem_2d_b200 as F  # noqa: E402
import oracle as O  # noqa: E402

CASES = [(seed, mesh, t, r, pmin, pmax)
         for seed, (mesh, t, r, pmin, pmax) in enumerate([
             (recipes.MESH_C, 1, 3, 1, 4), (recipes.MESH_C, 2, 2, 2, 6), (recipes.MESH_A, 0, 3, 1, 5), (recipes.MESH_A, 1, 2, 2, 4),
             (recipes.MESH_B, 0, 3, 2, 5), (recipes.MESH_B, 1, 1, 1, 3), (recipes.MESH_C, 0, 4, 3, 7), (recipes.MESH_A, 1, 3, 1, 3),
             (recipes.MESH_B, 1, 2, 3, 4), (recipes.MESH_C, 2, 3, 1, 2)], start=101)]


@pytest.mark.parametrize("seed,mesh,t_levels,rounds,pmin,pmax", CASES, ids=[f"seed{c[0]}" for c in CASES])
def test_random_hp_mesh_bit_identical(seed, mesh, t_levels, rounds, pmin, pmax):
    def build(api):
        return recipes.mesh_cfg4(api, t_levels=t_levels, rounds=rounds, pmin=pmin, pmax=pmax, seed=seed * 7919 + 13, mesh_file=mesh)
    do, df = O.Domain.from_mesh(build(recipes.api("oracle"))), F.Domain.from_mesh(build(recipes.api("product")))
    assert do.num_dofs == df.num_dofs
    if do.num_dofs == 0:
        pytest.skip("no DoFs on this mesh")
    nu, nv = 4 + seed % 5, 4 + (seed // 3) % 6
    glq = (F.gauss_quadrature_points(nu), F.gauss_quadrature_points(nv))
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq)
    for dedupe in (True, False):
        plan = F.Plan(df.view(), device=0, dedupe=dedupe)
        rows, cols, a, b = plan.assemble(glq)
        assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols), f"pattern differs (dedupe={dedupe})"
        assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(ref.a).view(np.uint64)), f"A differs (dedupe={dedupe})"
        assert np.array_equal(np.ascontiguousarray(b).view(np.uint64), np.ascontiguousarray(ref.b).view(np.uint64)), f"B differs (dedupe={dedupe})"
        assert plan.info["max_contrib"] <= 2
        hp = F.Plan(df.view(), device=-1, dedupe=dedupe)          # independent host construction of the pattern
        hr, hc = hp.pattern()
        assert np.array_equal(hr, rows) and np.array_equal(hc, cols)
        assert np.array_equal(hp.row_offsets(), plan.row_offsets())


@pytest.mark.parametrize("basis", [0, 1])
def test_throughput_shape_with_inter_layer_blocks_bit_identical(basis):
    """The 4-row tile shape (4 x 2 same-direction, 4 x 4 cross-direction tiles, 256- and 64-thread CTAs) on a mesh that has
    local-desc blocks of unequal BasisSpec lists (all four sub-blocks incl. V x U, RBS-scaled tables): the cfg-4 recipe at 3 T-levels
    (14 611 DoFs, 886 653 pairs, 622 blocks) against the oracle, bit for bit, unequal GLQ dims so that chunks end inside quadrature rows."""
    def build(api):
        return recipes.mesh_cfg4(api, t_levels=3, rounds=3, pmin=2, pmax=10)
    do, df = O.Domain.from_mesh(build(recipes.api("oracle"))), F.Domain.from_mesh(build(recipes.api("product")))
    glq = (F.gauss_quadrature_points(6), F.gauss_quadrature_points(7))
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq, basis=basis)
    for dedupe in (True, False):
        plan = F.Plan(df.view(), device=0, dedupe=dedupe)
        assert plan.info["tile_p"] == 4 and plan.info["n_pairs"] == 886653
        assert plan.check_work_items()["violations"] == 0
        rows, cols, a, b = plan.assemble(glq, basis=F.HierPoly if basis == 0 else F.HierMaxOrtho)
        assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
        assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(ref.a).view(np.uint64)), f"A differs (dedupe={dedupe})"
        assert np.array_equal(np.ascontiguousarray(b).view(np.uint64), np.ascontiguousarray(ref.b).view(np.uint64)), f"B differs (dedupe={dedupe})"


@pytest.mark.parametrize("nu,nv", [(5, 4), (9, 40)])
def test_throughput_shape_extreme_glq_shapes(nu, nv):
    """The same 14 611-DoF hp-mesh with GLQ shapes at both ends: 5 x 4 (a whole pack's rows fit one staged chunk; fewer points than the staging
    step holds) and 9 x 40 (one quadrature row of the widest pack no longer fits a ring buffer of the persistent integrator, so the launch
    falls back to k2_exact_kernel with 256-thread CTAs for every work item), bit for bit against the oracle."""
    def build(api):
        return recipes.mesh_cfg4(api, t_levels=3, rounds=3, pmin=2, pmax=10)
    do, df = O.Domain.from_mesh(build(recipes.api("oracle"))), F.Domain.from_mesh(build(recipes.api("product")))
    glq = (F.gauss_quadrature_points(nu), F.gauss_quadrature_points(nv))
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq, n_threads=16)
    plan = F.Plan(df.view(), device=0)
    assert plan.info["tile_p"] == 4
    rows, cols, a, b = plan.assemble(glq)
    assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(ref.a).view(np.uint64))
    assert np.array_equal(np.ascontiguousarray(b).view(np.uint64), np.ascontiguousarray(ref.b).view(np.uint64))
    # the second call re-uses the plan's scratch with the other tile-to-kernel routing
    glq2 = (F.gauss_quadrature_points(6), F.gauss_quadrature_points(6))
    ref2 = O.galerkin_sample_gep_hcurl(do, glq=glq2, n_threads=16)
    _, _, a2, b2 = plan.assemble(glq2)
    assert np.array_equal(np.ascontiguousarray(a2).view(np.uint64), np.ascontiguousarray(ref2.a).view(np.uint64))
    assert np.array_equal(np.ascontiguousarray(b2).view(np.uint64), np.ascontiguousarray(ref2.b).view(np.uint64))


@pytest.mark.parametrize("want_fold", [3, 1, 0])
def test_throughput_shape_folded_scales_bit_identical(want_fold, tmp_path):
    """The persistent integrator multiplies the quadrature weight instead of every product when a pass's scale is a power of two in every
    class of the plan (k2_ws_kernel FOLD): 3 = ratios and max(det) (dyadic Element sides: mesh a), 1 = ratios only (one 3 x 3 Element at two
    T-levels: square 1.5 / 2^k Elems, max(det) = 0.5625 / 4^k), 0 = neither (mesh c, 4.2 x 4.2: its refined sides carry rounding errors, the ratios of
    some classes are an ulp off 1).  High orders, no dedupe (the
    throughput tile shape needs 148 x 256 micro-tiles), 6 x 7 points, whole matrices against the oracle, bit for bit."""
    import json
    mesh3 = str(tmp_path / "mesh_3x3.json")
    with open(mesh3, "w") as f:
        json.dump({"Elements": [{"materials": [1.0, 0.0, 2.0, 0.0], "node_ids": [0, 1, 2, 3]}], "Nodes": [[0.0, 0.0], [3.0, 0.0], [0.0, 3.0], [3.0, 3.0]]}, f)

    def build(api):
        if want_fold == 3:
            m = api.Mesh.from_file(recipes.MESH_A)
            api.set_orders(m, 9, 9)
            m.global_h_refinement(api.href(recipes.T))
            m.h_refine_with_filter(lambda e: api.href(recipes.U) if e.id % 3 == 0 else api.href(recipes.V) if e.id % 3 == 1 else None)
            return m
        if want_fold == 1:
            m = api.Mesh.from_file(mesh3)
            api.set_orders(m, 12, 12)
            m.global_h_refinement(api.href(recipes.T)); m.global_h_refinement(api.href(recipes.T))
            return m
        return recipes.mesh_cfg4(api, t_levels=3, rounds=3, pmin=2, pmax=10)
    do, df = O.Domain.from_mesh(build(recipes.api("oracle"))), F.Domain.from_mesh(build(recipes.api("product")))
    glq = (F.gauss_quadrature_points(6), F.gauss_quadrature_points(7))
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq, n_threads=16)
    plan = F.Plan(df.view(), device=0, dedupe=False)
    assert plan.info["tile_p"] == 4 and plan.round_fill()["fold"] == want_fold
    rows, cols, a, b = plan.assemble(glq)
    assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(ref.a).view(np.uint64)), "A differs"
    assert np.array_equal(np.ascontiguousarray(b).view(np.uint64), np.ascontiguousarray(ref.b).view(np.uint64)), "B differs"
