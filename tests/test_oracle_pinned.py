"""Pins the CPU oracle (oracle/fem2d_oracle.cpp) on the reference's own fixtures, known-answer tests and doctests
(SURVEY.md section 8c).  Citations: /root/reference/<file>:<line>.  CPU only."""
import struct

import numpy as np
import pytest

import oracle as O
import recipes
from oracle import HRef, Mesh, Domain, T, U, V


@pytest.fixture(scope="module")
def slepc_gep():
    m = recipes.mesh_slepc(recipes.api("oracle"))
    d = Domain.from_mesh(m)
    return m, d, O.galerkin_sample_gep_hcurl(d, [8, 8])


def test_slepc_fixture_eigenpair(slepc_gep):
    """test_input/test_evec.dat + test_eval.dat are the SLEPc output of lib.rs:85-98: n = 624 and A x = lambda B x."""
    m, d, g = slepc_gep
    assert m.num_elems == 27 and d.num_dofs == 624 and len(g.rows) == 15512
    raw = open(recipes.GOLDEN + "/test_evec.dat", "rb").read()
    cid, n = struct.unpack(">ii", raw[:8])            # slepc_solve.rs:112-132
    assert cid == 1211214 and n == 624
    x = np.frombuffer(raw[8:], dtype=">f8").astype(np.float64)
    lam = struct.unpack(">d", open(recipes.GOLDEN + "/test_eval.dat", "rb").read())[0]
    assert lam == 1.4745880937038056
    A, B = g.dense()
    assert np.linalg.norm(A @ x - lam * (B @ x)) / np.linalg.norm(A @ x) < 5e-14
    assert abs((x @ A @ x) / (x @ B @ x) - lam) < 1e-13


def test_slepc_problem_eigenvalue(slepc_gep):
    import scipy.linalg as sl
    A, B = slepc_gep[2].dense()
    ev = sl.eigh(A, B, eigvals_only=True)
    assert abs(ev[np.argmin(abs(ev - 1.475))] - 1.4745880937) < 1e-9     # lib.rs:101-104


def test_nalg_problem_surrogate_eigenvalue():
    m = recipes.mesh_nalg(recipes.api("oracle"))
    d = Domain.from_mesh(m)
    g = O.galerkin_sample_gep_hcurl(d, [8, 8])
    assert d.num_dofs == 60 and len(g.rows) == 660
    assert abs(O.nalgebra_solve_surrogate(g, 2.64) - 2.6479657) < 1e-6     # lib.rs:62-65


def test_glq_20_point_table():
    """glq.rs:255-343 (X_20, W_20, X_20_SCALED), tolerance 1e-9 as in the reference."""
    X = [-0.993128599, -0.963971927, -0.912234428, -0.839116972, -0.746331906, -0.636053681, -0.510867002, -0.373706089, -0.227785851, -0.076526521]
    W = [0.017614007, 0.04060143, 0.062672048, 0.083276742, 0.10193012, 0.118194532, 0.131688638, 0.142096109, 0.149172986, 0.152753387]
    X20 = np.array(X + [-x for x in reversed(X)]); W20 = np.array(W + list(reversed(W)))
    p, w = O.gauss_quadrature_points(20)
    assert np.max(np.abs(p - X20)) < 1e-9 and np.max(np.abs(w - W20)) < 1e-9
    s = (0.5 - 0.25) / 2.0
    scaled = p * s + (0.5 + 0.25) / 2.0
    assert abs(s - 0.125) < 1e-14 and abs(scaled[0] - 0.250858925) < 1e-9 and abs(scaled[-1] - 0.499141075) < 1e-9


def test_glq_integral_doctest():
    p, w = O.gauss_quadrature_points(10)        # glq.rs:4-17: integral of u^2 v^2 = 4/9
    sol = 0.0
    for m in range(10):
        inner = 0.0
        for n in range(10):
            inner += p[m] ** 2 * p[n] ** 2 * w[n]
        sol += inner * w[m]
    assert abs(sol - 4.0 / 9.0) < 1e-12
    assert abs(p.sum()) < 1e-12
    assert [O.default_ngq(k) for k in (3, 4, 6, 8, 10)] == [16, 16, 32, 32, 64]    # basis.rs:172-177


def test_mesh_refinement_count_doctests():
    m = Mesh.unit(); assert m.num_elems == 1
    m.global_h_refinement(HRef(T)); assert m.num_elems == 5            # mesh.rs:703-710
    m.global_h_refinement(HRef(V)); assert m.num_elems == 13
    m = Mesh.unit()
    m.h_refine_elems([0], HRef(T)); assert m.num_elems == 5            # mesh.rs:740-751
    m.h_refine_elems([2, 3, 4], HRef(V)); assert m.num_elems == 11
    for bad in ([15], [0], [1, 1]):
        with pytest.raises(O.OracleError):
            m.h_refine_elems(bad, HRef(T))
    m = Mesh.unit()
    m.execute_h_refinements([(0, HRef(T))]); assert m.num_elems == 5   # mesh.rs:826-842
    m.execute_h_refinements([(1, HRef(U)), (1, HRef(V))]); assert m.num_elems == 9 and len(m.elem(1).children) == 4
    with pytest.raises(O.OracleError):
        m.execute_h_refinements([(2, HRef(T)), (0, HRef(T))])
    assert m.num_elems == 9


def test_descendant_ancestor_doctests():
    m = Mesh.unit()
    m.h_refine_elems([0], HRef(T)); m.h_refine_elems([1], HRef(T)); m.h_refine_elems([5], HRef(U))
    assert set(m.descendant_elems(1, True)) == {1, 5, 6, 7, 8, 9, 10}      # mesh.rs:453-468
    assert set(m.descendant_elems(1, False)) == {5, 6, 7, 8, 9, 10}
    assert m.ancestor_elems(10, True) == [10, 5, 1, 0]                      # mesh.rs:507-518
    assert m.ancestor_elems(10, False) == [5, 1, 0]


def test_sub_range_test_vectors():
    """h_refinement.rs:355-417"""
    import ctypes as C
    aniso = [[-1.0, 1.0, -1.0, 1.0], [-1.0, 1.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1.0], [0.0, 1.0, 0.0, 0.5], [0.0, 0.5, 0.0, 0.5]]
    iso = [[-1.0, 1.0, -1.0, 1.0], [-1.0, 0.0, -1.0, 0.0], [-0.5, 0.0, -1.0, -0.5], [-0.5, -0.25, -0.75, -0.5], [-0.375, -0.25, -0.625, -0.5]]
    N_, E_, S_, W_, SW, SE, NW, NE = 7, 5, 6, 4, 0, 1, 2, 3
    for locs, exp in (([N_, E_, S_, W_], aniso), ([SW, SE, NW, NE], iso)):
        r = np.array(exp[0])
        for k, loc in enumerate(locs):
            out = np.zeros(4)
            O.lib().orc_sub_range(loc, r.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
            assert np.max(np.abs(out - np.array(exp[k + 1]))) < 1e-14
            r = out


def test_basis_spec_count_doctests():
    m = Mesh.unit(); m.set_global_expansion_orders(2, 2)
    assert len(Domain.from_mesh(m).local_basis_specs(0)[0]) == 4           # domain.rs:245-251
    m = Mesh.unit(); m.set_global_expansion_orders(2, 2); m.global_h_refinement(HRef(T))
    d = Domain.from_mesh(m)
    dbs = d.descendant_basis_specs(0)
    assert len(dbs) == 4 and all(len(x[1][0]) == 8 for x in dbs)           # domain.rs:272-289
    m = Mesh.unit(); m.set_global_expansion_orders(2, 2); m.global_h_refinement(HRef(T)); m.h_refine_elems([1], HRef(T))
    d = Domain.from_mesh(m)
    abs_ = dict(d.ancestor_basis_specs(5))
    assert set(abs_) == {0, 1} and len(abs_[0][0]) == 0 and len(abs_[1][0]) == 4   # domain.rs:318-337


def test_mesh_a_file_geometry_and_neighbors():
    """mesh.rs:1850-1892 (mesh_from_file)"""
    m = Mesh.from_file(recipes.MESH_A)
    X = [[0.0, 1.0, 0.0, 1.0], [1.0, 2.0, 1.0, 2.0], [0.0, 1.0, 0.0, 1.0], [1.0, 2.0, 1.0, 2.0]]
    Y = [[0.0, 0.0, 0.5, 0.5], [0.0, 0.0, 0.5, 0.5], [0.5, 0.5, 1.0, 1.0], [0.5, 0.5, 1.0, 1.0]]
    NB = [[None, 2, None, 1], [None, 3, 0, None], [0, None, None, 3], [1, None, 2, None]]
    for e in range(4):
        pts = m.elem_points(e)
        for k in range(4):
            assert abs(pts[k][0] - X[e][k]) < 1e-14 and abs(pts[k][1] - Y[e][k]) < 1e-14
            if NB[e][k] is not None:
                act = m.edge(m.elem(e).edges[k])["active"]
                assert e in act and (act[0] + act[1] - e) == NB[e][k]


def test_proper_edge_order_recipe():
    """mesh.rs:1918-1948: incl. U(Some(1)) extended refinements; geometric invariants of every Elem."""
    m = recipes.mesh_edge_order(recipes.api("oracle"))
    assert m.num_elems == 297
    for e in range(m.num_elems):
        p = m.elem_points(e)
        assert p[0][0] < p[3][0] and p[0][1] < p[3][1]
        assert abs(p[0][1] - p[1][1]) < 1e-12 and abs(p[0][0] - p[2][0]) < 1e-12
        assert abs(p[3][1] - p[2][1]) < 1e-12 and abs(p[3][0] - p[1][0]) < 1e-12


def test_refinement_error_cases():
    """mesh.rs:1983-2082 (#[should_panic] tests)"""
    def mc():
        return Mesh.from_file(recipes.MESH_C)
    with pytest.raises(O.OracleError): mc().h_refine_elems([0, 1], HRef(T))
    m = mc(); m.h_refine_elems([0], HRef(T))
    with pytest.raises(O.OracleError): m.h_refine_elems([0], HRef(T))
    m = mc()
    with pytest.raises(O.OracleError):
        for _ in range(18):
            m.h_refine_with_filter(lambda e: HRef(T) if (not e.has_children and e.nodes[0] == 0) else None)
    with pytest.raises(O.OracleError): mc().p_refine_elems([0, 1], 1, 1)
    with pytest.raises(O.OracleError): mc().p_refine_elems([0, 0], 1, 1)
    m = mc(); m.set_global_expansion_orders(3, 3)
    with pytest.raises(O.OracleError): m.p_refine_elems([0], -3, 1)
    with pytest.raises(O.OracleError): m.p_refine_elems([0], 1, -3)
    with pytest.raises(O.OracleError): mc().p_refine_elems([0], 20, 0)
    with pytest.raises(O.OracleError): mc().p_refine_elems([0], 0, 20)


def test_create_domain_recipe_and_sizes():
    d = Domain.from_mesh(recipes.mesh_create_domain(recipes.api("oracle")))      # domain.rs:399-412
    assert d.mesh.num_elems == 36 and d.num_dofs == 1261
    d = Domain.from_mesh(recipes.mesh_readme(recipes.api("oracle")))             # BASELINE cfg 1
    g = O.galerkin_sample_gep_hcurl(d, [8, 8])
    assert d.mesh.num_elems == 28 and d.num_dofs == 600 and len(g.rows) == 13596


def test_error_order_and_threads_deterministic():
    d = Domain.from_mesh(recipes.mesh_nalg(recipes.api("oracle")))
    with pytest.raises(O.GalerkinSamplingError) as e:
        O.galerkin_sample_gep_hcurl(d, [3, 8])
    assert e.value.code == 3
    d = Domain.from_mesh(recipes.mesh_slepc(recipes.api("oracle")))
    g1 = O.galerkin_sample_gep_hcurl(d, [8, 8], n_threads=1)
    g4 = O.galerkin_sample_gep_hcurl(d, [8, 8], n_threads=4)
    assert np.array_equal(g1.a.view(np.uint64), g4.a.view(np.uint64)) and np.array_equal(g1.b.view(np.uint64), g4.b.view(np.uint64))


def test_anisotropic_mesh_reproduces_analytic_te_eigenvalues():
    """Physics check of the RBS inter-layer integrals: on the n-irregular U/V-refined 4.2 x 4.2 cavity with
    uniform p the lowest non-zero eigenvalues are (pi/4.2)^2 * {1, 1, 2, 4, 4, 5}."""
    import scipy.linalg as sl
    api = recipes.api("oracle")
    m = recipes.mesh_cfg4(api, t_levels=1, rounds=3, pmin=3, pmax=3)
    d = Domain.from_mesh(m)
    g = O.galerkin_sample_gep_hcurl(d, [8, 8])
    A, B = g.dense()
    ev = sl.eigh(A, B, eigvals_only=True)
    ev = ev[ev > 1e-6][:6] / (np.pi / 4.2) ** 2
    assert np.allclose(ev, [1, 1, 2, 4, 4, 5], rtol=1e-2) and np.allclose(ev[:3], [1, 1, 2], rtol=5e-4), ev
