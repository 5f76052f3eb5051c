"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): CSR structure bit-exact; EXACT mode values bit-identical to the oracle (so the literal
1e-12 relative / 1e-14 absolute tolerance holds trivially).  Re-ordered modes are judged with the scale-aware floor."""
import numpy as np
import pytest

import recipes

pytestmark = pytest.mark.gpu

import fem_2d_b200 as F  # noqa: E402
import oracle as O  # noqa: E402


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)


def _assert_bit_identical(got, ref, what):
    gb, rb = _bits(got), _bits(ref)
    if not np.array_equal(gb, rb):
        bad = np.nonzero(gb != rb)[0]
        k = bad[0]
        raise AssertionError(f"{what}: {len(bad)}/{len(gb)} entries differ bitwise; first at {k}: got {got[k]!r} ref {ref[k]!r}")


def _glq(nu, nv):
    # one node set shared by both sides: nodes/weights are inputs of the path
    return (F.gauss_quadrature_points(nu), F.gauss_quadrature_points(nv))


def _run_pair(name, nu, nv, basis=0, dedupe=True, mode=F.MODE_EXACT):
    mo, mf = recipes.build_pair(name)
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    glq = _glq(nu, nv)
    ref = O.galerkin_sample_gep_hcurl(do, basis=basis, glq=glq)
    plan = F.Plan(df.view(), device=0, dedupe=dedupe)
    rows, cols, a, b = plan.assemble(glq, basis=F.HierPoly if basis == 0 else F.HierMaxOrtho, mode=mode)
    return ref, plan, rows, cols, a, b


@pytest.mark.parametrize("name", sorted(recipes.RECIPES))
@pytest.mark.parametrize("dedupe", [True, False])
def test_exact_bit_identical(name, dedupe):
    ref, plan, rows, cols, a, b = _run_pair(name, 8, 8, dedupe=dedupe)
    assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
    _assert_bit_identical(a, ref.a, f"A[{name}]")
    _assert_bit_identical(b, ref.b, f"B[{name}]")
    assert plan.info["max_contrib"] <= 2


@pytest.mark.parametrize("nu,nv", [(4, 4), (5, 9), (12, 12), (7, 16), (32, 32)])
def test_exact_glq_shapes(nu, nv):
    # odd counts put a node at 0 (signed-zero paths), nu != nv exercises the (m outer, n inner) order
    for name in ("slepc", "cfg4_small"):
        ref, plan, rows, cols, a, b = _run_pair(name, nu, nv)
        _assert_bit_identical(a, ref.a, f"A[{name},{nu}x{nv}]")
        _assert_bit_identical(b, ref.b, f"B[{name},{nu}x{nv}]")


def test_exact_max_ortho():
    for name in ("slepc", "readme", "cfg4_small"):
        ref, plan, rows, cols, a, b = _run_pair(name, 8, 8, basis=1)
        _assert_bit_identical(a, ref.a, f"A[{name},HierMaxOrtho]")
        _assert_bit_identical(b, ref.b, f"B[{name},HierMaxOrtho]")


def test_device_pattern_matches_host_pattern():
    for name in ("slepc", "edge_order", "cfg4_small"):
        _, mf = recipes.build_pair(name)
        df = F.Domain.from_mesh(mf)
        pd = F.Plan(df.view(), device=0)
        ph = F.Plan(df.view(), device=-1)
        rd, cd = pd.pattern(); rh, ch = ph.pattern()
        assert np.array_equal(rd, rh) and np.array_equal(cd, ch)
        for k in ("nnz_upper", "n_pairs", "n_multi", "max_contrib", "n_extra"):
            assert pd.info[k] == ph.info[k], k
        assert np.array_equal(pd.row_blocks(3), ph.row_blocks(3))


def _pair_from(build):
    return build(recipes.api("oracle")), build(recipes.api("product"))


def test_exact_at_the_limits():
    """Maximum sizes of the path: MAX_POLYNOMIAL_ORDER = 20 (mesh.rs:42; 1 680 functions per Elem, multi-round work items and the widest
    slab rows), 128 GLQ points per axis (= default_ngq(20), basis.rs:172-177; many staging chunks) and the reference's default GLQ
    (glq_grid_dim = None)."""
    def order20(api):
        m = api.Mesh.from_file(recipes.MESH_C)
        api.set_orders(m, 20, 20)
        return m
    mo, mf = _pair_from(order20)
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    glq = _glq(6, 5)
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq)
    rows, cols, a, b = F.Plan(df.view(), device=0).assemble(glq)
    assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
    _assert_bit_identical(a, ref.a, "A[order 20]"); _assert_bit_identical(b, ref.b, "B[order 20]")

    mo, mf = recipes.build_pair("nalg")
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    glq = _glq(128, 128)
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq)
    for dedupe in (True, False):
        rows, cols, a, b = F.Plan(df.view(), device=0, dedupe=dedupe).assemble(glq)
        _assert_bit_identical(a, ref.a, "A[128x128]"); _assert_bit_identical(b, ref.b, "B[128x128]")

    # glq_grid_dim = None -> default_ngq(max order) points per axis (galerkin.rs:66-67, basis.rs:83-90): order 3 -> 16 x 16
    assert F.default_ngq(3) == 16 and F.default_ngq(20) == 128
    gep = F.galerkin_sample_gep_hcurl(df, None, device=0)
    ref = O.galerkin_sample_gep_hcurl(do, glq=_glq(16, 16))
    _assert_bit_identical(gep.a.values, ref.a, "A[default GLQ]"); _assert_bit_identical(gep.b.values, ref.b, "B[default GLQ]")


def test_repeatable_across_calls_streams_and_launch_modes():
    """Deterministic by construction (fixed source order per slot, no atomics): repeated calls, calls on a side stream, calls with
    the per-phase events on (kernels not overlapped) and off (programmatic dependent launch) all give the same bits."""
    import torch
    df = F.Domain.from_mesh(recipes.build_pair("cfg4_small")[1])
    glq = _glq(12, 12)
    plan = F.Plan(df.view(), device=0)
    outs = []
    side = torch.cuda.Stream()
    for k in range(6):
        a = torch.full((plan.nnz,), float("nan"), dtype=torch.float64, device="cuda:0"); b = torch.full_like(a, float("nan"))
        plan.set_phase_timing(k % 2 == 1)
        stream = side if k >= 3 else torch.cuda.current_stream()
        stream.wait_stream(torch.cuda.current_stream())
        plan.assemble_device(glq, a.data_ptr(), b.data_ptr(), stream=stream.cuda_stream)
        stream.synchronize()
        outs.append((a, b))
    for a, b in outs[1:]:
        assert torch.equal(a.view(torch.int64), outs[0][0].view(torch.int64)) and torch.equal(b.view(torch.int64), outs[0][1].view(torch.int64))
    assert not torch.isnan(outs[0][0]).any() and not torch.isnan(outs[0][1]).any()
    t = plan.last_timing()
    assert t["launches"] >= 3 and t["total_ms"] > 0


def test_device_row_offsets_and_sliced_rows():
    """Host `rows` are expanded from the device CSR row offsets (fem2d_plan_row_offsets), also for slices that start mid-row."""
    df = F.Domain.from_mesh(recipes.build_pair("cfg4_small")[1])
    plan = F.Plan(df.view(), device=0)
    rows, cols = plan.pattern()
    rp = plan.row_offsets()
    assert np.array_equal(np.repeat(np.arange(df.num_dofs, dtype=np.uint32), np.diff(rp).astype(np.int64)), rows)
    xfer = plan.pattern_transfer_info()   # the host calls move row offsets + column runs instead of rows[] / cols[]
    assert xfer["row_offset_bytes"] == 4 * (df.num_dofs + 1) and df.num_dofs <= xfer["col_runs"] <= plan.nnz
    assert xfer["plain_bytes"] == 8 * plan.nnz
    glq = _glq(8, 8)
    _, _, a_full, b_full = plan.assemble(glq)
    n = plan.nnz
    ranges = [(3, n // 3 + 1), (n // 2 + 5, n - 2)]
    m = sum(e - b for b, e in ranges)
    r2 = np.zeros(m, dtype=np.uint32); c2 = np.zeros(m, dtype=np.uint32); a2 = np.zeros(m); b2 = np.zeros(m)
    plan.assemble_ranges_into(glq, ranges, a2.ctypes.data, b2.ctypes.data, r2.ctypes.data, c2.ctypes.data)
    sel = np.concatenate([np.arange(b, e) for b, e in ranges])
    assert np.array_equal(r2, rows[sel]) and np.array_equal(c2, cols[sel])
    _assert_bit_identical(a2, a_full[sel], "A slices"); _assert_bit_identical(b2, b_full[sel], "B slices")


def test_reference_call_and_errors():
    _, mf = recipes.build_pair("nalg")
    df = F.Domain.from_mesh(mf)
    gep = F.galerkin_sample_gep_hcurl(df, [8, 8])   # lib.rs:57-59
    assert gep.a.dimension == 60 and len(gep.a.rows) == 660
    sol = F.nalgebra_solve_gep(gep, 2.64)           # lib.rs:62-66
    assert abs(sol.value - 2.6479657) < 1e-6
    assert len(sol.vector) == df.num_dofs
    with pytest.raises(F.GalerkinSamplingError) as e:
        F.galerkin_sample_gep_hcurl(df, [3, 8])
    assert e.value.kind == F.GalerkinSamplingError.InvalidGLQSettings
    with pytest.raises(F.GalerkinSamplingError) as e:
        F.galerkin_sample_gep_hcurl(F.Domain.blank(F.ContinuityCondition.HDiv), [8, 8])
    assert e.value.kind == F.GalerkinSamplingError.WrongContinuityCondition
    with pytest.raises(F.GalerkinSamplingError) as e:
        F.galerkin_sample_gep_hcurl(F.Domain.blank(F.ContinuityCondition.HCurl), [8, 8])
    assert e.value.kind == F.GalerkinSamplingError.EmptyDOFSet
    # default GLQ (None): basis.rs:172-177 -> 16 points per axis for order 3... (4*3=12 -> 16)
    gep2 = F.galerkin_sample_gep_hcurl(df, None)
    ref = O.galerkin_sample_gep_hcurl(O.Domain.from_mesh(recipes.build_pair("nalg")[0]), glq=(F.gauss_quadrature_points(16), F.gauss_quadrature_points(16)))
    _assert_bit_identical(gep2.a.values, ref.a, "A default glq")


def test_slepc_fixture_residual_with_gpu_matrices():
    """The reference's stored eigenpair (test_input/test_evec.dat, test_eval.dat) must satisfy A x = lambda B x with the GPU matrices."""
    import struct
    _, mf = recipes.build_pair("slepc")
    df = F.Domain.from_mesh(mf)
    assert df.num_dofs == 624
    gep = F.galerkin_sample_gep_hcurl(df, [8, 8])
    raw = open(recipes.GOLDEN + "/test_evec.dat", "rb").read()
    cid, n = struct.unpack(">ii", raw[:8])
    assert cid == 1211214 and n == 624
    x = np.frombuffer(raw[8:], dtype=">f8").astype(np.float64)
    lam = struct.unpack(">d", open(recipes.GOLDEN + "/test_eval.dat", "rb").read())[0]
    A, B = gep.to_nalgebra_dense_mats()
    r = A @ x - lam * (B @ x)
    assert np.linalg.norm(r) / np.linalg.norm(A @ x) < 1e-12
    assert abs((x @ A @ x) / (x @ B @ x) - 1.4745880937) < 1e-9   # lib.rs:104


def _assert_scale_aware(got, ref, what, rel=1e-12, floor=1e-14):
    """Re-ordered modes: |got - ref| <= rel*|ref| + floor*max|M| (the literal 1e-14 absolute floor is below
    the round-off noise of the reference's own parity cancellations, ~1e-16*max|M|)."""
    scale = np.max(np.abs(ref))
    err = np.abs(got - ref)
    tol = rel * np.abs(ref) + floor * scale
    bad = np.nonzero(err > tol)[0]
    assert len(bad) == 0, f"{what}: {len(bad)}/{len(ref)} entries out of tolerance, worst {np.max(err / tol):.3g}x at {bad[np.argmax((err / tol)[bad])]}"


@pytest.mark.parametrize("mode", ["sumfact", "dmma"])
@pytest.mark.parametrize("name,nu,nv", [("slepc", 8, 8), ("readme", 8, 8), ("edge_order", 5, 9), ("cfg4_small", 12, 12), ("cfg3_small", 8, 8),
                                        ("create_domain", 16, 16)])
def test_reordered_modes_scale_aware(mode, name, nu, nv):
    m = {"sumfact": F.MODE_SUMFACT, "dmma": F.MODE_DMMA}[mode]
    for dedupe in (True, False):
        ref, plan, rows, cols, a, b = _run_pair(name, nu, nv, dedupe=dedupe, mode=m)
        assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
        _assert_scale_aware(a, ref.a, f"A[{name},{mode}]")
        _assert_scale_aware(b, ref.b, f"B[{name},{mode}]")
        # measured: <= 8e-16 * max|M| absolute, <= 6e-15 relative on entries above 1e-6 * max|M| (well inside the bar); which
        # parity-cancellation entries come out as exact zeros differs between summation orders, so that is not asserted


def test_reordered_modes_same_eigenvalue():
    """'matching eigenvalues from the same downstream nalgebra solve' (north_star) also for the re-ordered modes."""
    _, mf = recipes.build_pair("nalg")
    df = F.Domain.from_mesh(mf)
    for m in (F.MODE_SUMFACT, F.MODE_DMMA):
        gep = F.galerkin_sample_gep_hcurl(df, [8, 8], mode=m)
        assert abs(F.nalgebra_solve_gep(gep, 2.64).value - 2.6479657) < 1e-6


def test_swapped_integrals_and_slices():
    _, mf = recipes.build_pair("readme")
    df = F.Domain.from_mesh(mf)
    glq = _glq(8, 8)
    plan = F.Plan(df.view(), device=0)
    r, c, a, b = plan.assemble(glq)
    r2, c2, a2, b2 = plan.assemble(glq, a=F.L2Inner, b=F.CurlCurl)
    _assert_bit_identical(a2, b, "swapped A")
    _assert_bit_identical(b2, a, "swapped B")
    # row-block slices through the device entry point reproduce the full result
    import torch
    da = torch.full((plan.nnz,), float("nan"), dtype=torch.float64, device="cuda:0")
    db = torch.full((plan.nnz,), float("nan"), dtype=torch.float64, device="cuda:0")
    bounds = plan.row_blocks(3)
    for k in range(3):
        plan.assemble_device(glq, da.data_ptr(), db.data_ptr(), slot_begin=int(bounds[k]), slot_end=int(bounds[k + 1]),
                             stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    _assert_bit_identical(da.cpu().numpy(), a, "sliced A")
    _assert_bit_identical(db.cpu().numpy(), b, "sliced B")


def test_trim_cache_returns_memory_and_the_library_keeps_working():
    """fem2d_trim_cache hands the parked device blocks / pinned staging buffers back to the driver; the next plan allocates afresh and the
    results do not change."""
    import torch
    mo, mf = recipes.build_pair("cfg4_small")
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    glq = _glq(8, 8)
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq)
    for _ in range(2):
        rows, cols, a, b = F.Plan(df.view(), device=0).assemble(glq)
        _assert_bit_identical(a, ref.a, "A"); _assert_bit_identical(b, ref.b, "B")
        torch.cuda.synchronize()
        free0 = torch.cuda.mem_get_info()[0]
        F.trim_cache()
        assert torch.cuda.mem_get_info()[0] >= free0
