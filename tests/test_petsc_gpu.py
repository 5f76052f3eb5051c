"""PETSc AIJ emission from the device arrays against an oracle-side restatement of sparse_matrix.rs:184-264, byte for byte,
on BASELINE cfg 5 (the slepc_problem domain) and cfg 4 (anisotropic hp-mesh); and against the host writer of the host mirror."""
import numpy as np
import pytest

import recipes

pytestmark = pytest.mark.gpu

import fem_2d_b200 as F  # noqa: E402
import oracle as O  # noqa: E402


@pytest.mark.parametrize("name,g", [("slepc", 8), ("readme", 8), ("cfg4_full", 12)])
def test_device_petsc_image_matches_oracle_writer(name, g, tmp_path):
    import torch
    mf = recipes.mesh_cfg4(recipes.api("product")) if name == "cfg4_full" else recipes.RECIPES[name](recipes.api("product"))
    df = F.Domain.from_mesh(mf)
    glq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
    plan = F.Plan(df.view(), device=0)
    da = torch.empty(plan.nnz, dtype=torch.float64, device="cuda:0"); db = torch.empty_like(da)
    plan.assemble_device(glq, da.data_ptr(), db.data_ptr())
    torch.cuda.synchronize()
    rows, cols = plan.pattern()
    for d_vals in (da, db):
        ref = O.petsc_aij_bytes(plan.n_dofs, rows, cols, d_vals.cpu().numpy())
        img = plan.petsc_aij_image(d_vals.data_ptr())
        assert img.tobytes() == ref
    # file form + the host mirror's writer (SparseMatrix::print_to_petsc_binary_file of the Python face)
    p1, p2 = str(tmp_path / "dev_a.dat"), str(tmp_path / "host_a.dat")
    plan.write_petsc_aij(da.data_ptr(), p1)
    F.SparseMatrix(plan.n_dofs, rows, cols, da.cpu().numpy()).print_to_petsc_binary_file(p2)
    b1, b2 = open(p1, "rb").read(), open(p2, "rb").read()
    assert b1 == b2 == O.petsc_aij_bytes(plan.n_dofs, rows, cols, da.cpu().numpy())
    assert b1[:4] == bytes([0x00, 0x12, 0x7B, 0x50])
    nf = 2 * plan.nnz - plan.n_dofs            # every DoF has its diagonal entry
    assert len(b1) == 16 + 4 * plan.n_dofs + 12 * nf
