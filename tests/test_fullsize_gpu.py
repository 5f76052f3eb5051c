"""Full-size checks at BASELINE.json configs[2] (mesh_a, Orders(6,6), 6x global T: 1,178,112 DoFs), where the CPU oracle cannot
run the whole assembly in test time: size-independent properties plus oracle spot checks of individual Elems."""
import numpy as np
import pytest

import recipes

pytestmark = pytest.mark.gpu

import fem_2d_b200 as F  # noqa: E402
import oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def full():
    import torch
    mf = recipes.mesh_cfg3(recipes.api("product"), levels=6, order=6)
    df = F.Domain.from_mesh(mf)
    glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
    plan = F.Plan(df.view(), device=0, dedupe=True)
    a = torch.empty(plan.nnz, dtype=torch.float64, device="cuda:0"); b = torch.empty_like(a)
    plan.assemble_device(glq, a.data_ptr(), b.data_ptr())
    torch.cuda.synchronize()
    return dict(mesh=mf, domain=df, glq=glq, plan=plan, a=a, b=b)


def test_sizes_match_closed_form(full):
    info = full["plan"].info
    assert full["domain"].num_dofs == 1178112            # SURVEY.md section 8 table, cfg 3
    assert info["nnz_upper"] == 57557904 and info["n_pairs"] == 58240656
    assert info["max_contrib"] == 2 and info["n_extra"] == 58240656 - 57557904


def test_source_map_packs_and_keeps_plain_chunks(full):
    """The scatter kernel reads a packed source map: most 64-slot chunks as 16-bit offsets, the chunks holding multi-contribution slots
    (or spanning more than 2^16 value entries) in plain 32-bit form.  Both forms must occur at this size, so the bit-identity checks
    in this file cover both code paths of the kernel."""
    smi = full["plan"].source_map_info()
    n_chunks = -(-full["plan"].nnz // smi["chunk_slots"])
    assert 0 < smi["plain_chunks"] < 0.1 * n_chunks
    assert smi["map_bytes"] < 0.6 * smi["plain_bytes"]
    nd = F.Plan(full["domain"].view(), device=0, dedupe=False).source_map_info()
    assert smi["plain_chunks"] <= nd["plain_chunks"] < 0.1 * n_chunks      # dedupe off: edge rows reach across distant value tiles


def test_host_pattern_expansion_at_full_size(full):
    """rows[] / cols[] of the host-output calls are expanded on host threads from the CSR row offsets and the column runs: at full
    size (multi-threaded, streaming stores, 57.6 M slots) they must equal the device pattern."""
    plan = full["plan"]
    rows, cols = plan.pattern()                     # plain D2H of the device arrays
    xfer = plan.pattern_transfer_info()
    assert xfer["row_offset_bytes"] + xfer["col_run_bytes"] < 0.25 * xfer["plain_bytes"]
    n = plan.nnz
    r2 = np.zeros(n, dtype=np.uint32); c2 = np.zeros(n, dtype=np.uint32); a2 = np.zeros(n); b2 = np.zeros(n)
    plan.assemble_ranges_into(full["glq"], [(0, n)], a2.ctypes.data, b2.ctypes.data, r2.ctypes.data, c2.ctypes.data)
    assert np.array_equal(r2, rows) and np.array_equal(c2, cols)
    assert np.array_equal(a2.view(np.uint64), full["a"].cpu().numpy().view(np.uint64))


def test_pattern_is_sorted_unique_upper_triangular(full):
    rows, cols = full["plan"].pattern()
    assert np.all(rows <= cols)
    keys = rows.astype(np.int64) << 32 | cols.astype(np.int64)
    assert np.all(np.diff(keys) > 0)
    assert rows[0] == 0 and cols.max() == full["domain"].num_dofs - 1
    # every DoF has its diagonal entry
    assert np.count_nonzero(rows == cols) == full["domain"].num_dofs


def test_dedupe_off_is_bit_identical(full):
    import torch
    plan = F.Plan(full["domain"].view(), device=0, dedupe=False)
    assert plan.info["n_classes"] == plan.info["n_blocks"] == 16384
    a = torch.empty_like(full["a"]); b = torch.empty_like(full["b"])
    plan.assemble_device(full["glq"], a.data_ptr(), b.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int64), full["a"].view(torch.int64))
    assert torch.equal(b.view(torch.int64), full["b"].view(torch.int64))


def test_material_scaling_is_exact(full):
    """A = (1/mu) * (...), B = eps * (...): scaling by powers of two is exact in IEEE arithmetic, whatever the mesh size."""
    import torch
    df = F.Domain.from_mesh(recipes.mesh_cfg3(recipes.api("product"), levels=6, order=6))
    v = df.view()
    v.element_eps_re *= 2.0
    v.element_mu_re *= 4.0
    plan = F.Plan(v, device=0)
    a = torch.empty_like(full["a"]); b = torch.empty_like(full["b"])
    plan.assemble_device(full["glq"], a.data_ptr(), b.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal((a * 4.0).view(torch.int64), full["a"].view(torch.int64))
    assert torch.equal((b * 0.5).view(torch.int64), full["b"].view(torch.int64))


def test_oracle_spot_check_of_single_elems(full):
    """Per-Elem matrices of a few leaves (corner, edge, interior, last) from the oracle vs the GPU values at the same keys.
    Keys between two Elem-type DoFs of one leaf receive exactly one contribution, so they must match bit for bit; all keys of
    a leaf must match once the neighbours' contributions are added -- checked through the full sum for an interior leaf."""
    mo = recipes.mesh_cfg3(recipes.api("oracle"), levels=6, order=6)
    do = O.Domain.from_mesh(mo)
    n_leaf0 = 21844 - 16384
    ids = [n_leaf0, n_leaf0 + 1, n_leaf0 + 777, n_leaf0 + 8191, 21843]
    el, r, c, a, b = O.assemble_elems(do, ids, full["glq"])
    rows, cols = full["plan"].pattern()
    keys = rows.astype(np.int64) << 32 | cols.astype(np.int64)
    slot = np.searchsorted(keys, r.astype(np.int64) << 32 | c.astype(np.int64))
    assert np.array_equal(keys[slot], r.astype(np.int64) << 32 | c.astype(np.int64))     # every oracle key exists in the pattern
    ga = full["a"].cpu().numpy()[slot]; gb = full["b"].cpu().numpy()[slot]
    n_elem_type = 16384 * 60                                                             # Elem-type DoFs are numbered first (domain.rs:83-96)
    single = (r < n_elem_type) & (c < n_elem_type)
    assert single.sum() == len(ids) * 60 * 61 // 2
    assert np.array_equal(ga[single].view(np.uint64), a[single].view(np.uint64))
    assert np.array_equal(gb[single].view(np.uint64), b[single].view(np.uint64))
    # shared (edge-type) keys of one interior leaf: add the contributions of its four edge neighbours (<= 2 terms per key, so the
    # order of the IEEE sum is immaterial) and compare every key of that leaf bit for bit
    centre = None
    for cand in range(n_leaf0 + 8191, 21844):      # first leaf from the middle of the id range whose four edges are all interior
        acts = [mo.edge(e)["active"] for e in mo.elem(cand).edges]
        if all(cand in a for a in acts):
            centre = cand
            break
    assert centre is not None
    neigh = {a[0] + a[1] - centre for a in acts}
    assert len(neigh) == 4
    el, r, c, a, b = O.assemble_elems(do, [centre] + sorted(neigh), full["glq"])
    k_all = r.astype(np.int64) << 32 | c.astype(np.int64)
    mine = el == centre
    tot_a, tot_b = {}, {}
    for k, va, vb in zip(k_all[mine], a[mine], b[mine]):
        tot_a[k] = va; tot_b[k] = vb
    n_two = 0
    for k, va, vb in zip(k_all[~mine], a[~mine], b[~mine]):
        if k in tot_a:
            tot_a[k] = tot_a[k] + va; tot_b[k] = tot_b[k] + vb; n_two += 1
    assert n_two == 4 * 21                                                               # 6 edge functions per shared edge: 6*7/2 keys
    kk = np.array(sorted(tot_a), dtype=np.int64)
    slot = np.searchsorted(keys, kk)
    assert np.array_equal(keys[slot], kk)
    ref_a = np.array([tot_a[k] for k in kk]); ref_b = np.array([tot_b[k] for k in kk])
    assert np.array_equal(full["a"].cpu().numpy()[slot].view(np.uint64), ref_a.view(np.uint64))
    assert np.array_equal(full["b"].cpu().numpy()[slot].view(np.uint64), ref_b.view(np.uint64))


def test_row_block_slices_with_restricted_integrator():
    """Multi-GPU sharding on one device: every row block is assembled on its own (the integrator only runs the micro-tiles
    that block reads); the blocks together reproduce the full-range result bit for bit.  Dedupe off = general case."""
    import torch
    df = F.Domain.from_mesh(recipes.mesh_cfg3(recipes.api("product"), levels=4, order=6))
    glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
    plan = F.Plan(df.view(), device=0, dedupe=False)
    ref_a = torch.empty(plan.nnz, dtype=torch.float64, device="cuda:0"); ref_b = torch.empty_like(ref_a)
    plan.assemble_device(glq, ref_a.data_ptr(), ref_b.data_ptr())
    a = torch.full_like(ref_a, float("nan")); b = torch.full_like(ref_b, float("nan"))
    world = 4
    bounds = plan.row_blocks(world)
    total_tiles = 1024 * 363                 # 1024 leaves, 363 micro-tiles (4x2 same-, 4x4 cross-direction) per interior leaf class (fewer on the boundary)
    needed = []
    for r in range(world):
        plan.assemble_device(glq, a.data_ptr(), b.data_ptr(), slot_begin=int(bounds[r]), slot_end=int(bounds[r + 1]))
        needed.append(plan.refresh_info()["range_tiles_needed"])
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int64), ref_a.view(torch.int64))
    assert torch.equal(b.view(torch.int64), ref_b.view(torch.int64))
    assert all(0 < n < 0.6 * total_tiles for n in needed), needed       # each rank integrates a fraction of the tiles
    assert sum(needed) < 1.6 * total_tiles, needed                        # little redundancy across ranks


def test_split_partition_balances_the_integrator():
    """fem2d_plan_row_blocks_split: a rank's Elem-type rows and its edge-type rows in one multi-range call.  All ranks together
    reproduce the full result bit for bit and need a balanced share of the micro-tiles (dedupe off = general case)."""
    import torch
    df = F.Domain.from_mesh(recipes.mesh_cfg3(recipes.api("product"), levels=4, order=6))
    glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
    plan = F.Plan(df.view(), device=0, dedupe=False)
    ref_a = torch.empty(plan.nnz, dtype=torch.float64, device="cuda:0"); ref_b = torch.empty_like(ref_a)
    plan.assemble_device(glq, ref_a.data_ptr(), ref_b.data_ptr())
    a = torch.full_like(ref_a, float("nan")); b = torch.full_like(ref_b, float("nan"))
    world = 4
    b1, b2 = plan.row_blocks_split(world)
    assert b1[0] == 0 and b1[-1] == b2[0] and b2[-1] == plan.nnz
    rows, _ = plan.pattern()
    assert rows[int(b2[0]) - 1] < 1024 * 60 <= rows[int(b2[0])]           # the split sits at the first edge-type DoF row
    needed = []
    for r in range(world):
        plan.assemble_device_ranges(glq, a.data_ptr(), b.data_ptr(), [(int(b1[r]), int(b1[r + 1])), (int(b2[r]), int(b2[r + 1]))])
        needed.append(plan.refresh_info()["range_tiles_needed"])
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int64), ref_a.view(torch.int64))
    assert torch.equal(b.view(torch.int64), ref_b.view(torch.int64))
    total_tiles = 1024 * 363
    assert max(needed) < 1.35 * min(needed), needed                      # balanced
    assert sum(needed) < 1.35 * total_tiles, needed                      # little redundancy


def test_split_partition_on_the_anisotropic_hp_mesh():
    """BASELINE.json configs[3] (57 825 DoFs, 2 620 classes incl. local-desc blocks, GLQ 12 x 12) sharded over 3 ranks: the restricted
    integrator (multi-range work items that stage only the functions their tiles touch, 4 x 4 cross-direction tiles, V x U sub-blocks)
    reproduces the full-range result bit for bit, with both dedupe settings."""
    import torch
    df = F.Domain.from_mesh(recipes.mesh_cfg4(recipes.api("product")))
    glq = (F.gauss_quadrature_points(12), F.gauss_quadrature_points(12))
    for dedupe in (True, False):
        plan = F.Plan(df.view(), device=0, dedupe=dedupe)
        assert plan.n_dofs == 57825 and plan.nnz == 3505664 and plan.info["tile_p"] == 4
        chk = plan.check_work_items()
        assert chk["violations"] == 0
        ref_a = torch.empty(plan.nnz, dtype=torch.float64, device="cuda:0"); ref_b = torch.empty_like(ref_a)
        plan.assemble_device(glq, ref_a.data_ptr(), ref_b.data_ptr())
        a = torch.full_like(ref_a, float("nan")); b = torch.full_like(ref_b, float("nan"))
        world = 3
        b1, b2 = plan.row_blocks_split(world)
        needed = []
        for r in range(world):
            plan.assemble_device_ranges(glq, a.data_ptr(), b.data_ptr(), [(int(b1[r]), int(b1[r + 1])), (int(b2[r]), int(b2[r + 1]))])
            needed.append(plan.refresh_info()["range_tiles_needed"])
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int64), ref_a.view(torch.int64))
        assert torch.equal(b.view(torch.int64), ref_b.view(torch.int64))
        assert all(0 < n < chk["tiles"] for n in needed), (needed, chk)     # every rank ran a restricted item list
