"""BASELINE.json configs[4]: assembly followed by xy_fields E-field evaluation on UniformFieldSpace [16,16] from the stored
test_evec.dat eigenvector (fields.rs:63-127), GPU vs oracle, bit for bit."""
import struct

import numpy as np
import pytest

import recipes

pytestmark = pytest.mark.gpu

import fem_2d_b200 as F  # noqa: E402
import oracle as O  # noqa: E402


def _evec():
    raw = open(recipes.GOLDEN + "/test_evec.dat", "rb").read()
    return np.frombuffer(raw[8:], dtype=">f8").astype(np.float64)


def _check(name, density, solution=None, basis=0):
    mo, mf = recipes.build_pair(name)
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    sol = solution if solution is not None else np.cos(np.arange(df.num_dofs) * 0.37) + 0.25
    ids, x, y = O.xy_fields(do, [density, density], sol, basis=basis)
    ufs = F.UniformFieldSpace(df, [density, density])
    xn, yn = ufs.xy_fields("E", sol, basis=F.HierPoly if basis == 0 else F.HierMaxOrtho)
    assert (xn, yn) == ("E_x", "E_y")
    assert sorted(ufs.quantities[xn]) == ids.tolist()
    for k, e in enumerate(ids.tolist()):
        assert np.array_equal(ufs.quantities[xn][e].view(np.uint64), x[k].view(np.uint64)), (name, e, "x")
        assert np.array_equal(ufs.quantities[yn][e].view(np.uint64), y[k].view(np.uint64)), (name, e, "y")
    return ufs, ids


def test_cfg5_fields_from_stored_eigenvector():
    x = _evec()
    ufs, ids = _check("slepc", 16, solution=x / np.sqrt(np.sum(x ** 2)))      # normalized_eigenvector (linalg.rs:93-96), lib.rs:108-110
    assert len(ids) == 21                                                       # 21 leaf Elems (BASELINE.md cfg 5)
    ufs.expression_2arg(["E_x", "E_y"], "E_mag", lambda ex, ey: np.sqrt(ex ** 2 + ey ** 2))   # lib.rs:111-115
    assert set(ufs.quantities["E_mag"]) == set(ids.tolist())
    assert max(np.max(v) for v in ufs.quantities["E_mag"].values()) > 0.0


@pytest.mark.parametrize("name,density", [("readme", 8), ("edge_order", 5), ("cfg4_small", 16), ("create_domain", 10)])
def test_fields_match_oracle_on_refined_meshes(name, density):
    _check(name, density)


def test_fields_max_ortho_and_errors():
    _check("readme", 8, basis=1)
    _, mf = recipes.build_pair("nalg")
    df = F.Domain.from_mesh(mf)
    with pytest.raises(F.UniformFieldError):
        F.UniformFieldSpace(df, [8, 8]).xy_fields("E", np.zeros(df.num_dofs + 1))    # MismatchedSolutionSize (fields.rs:68-72)
    with pytest.raises(F.UniformFieldError):
        F.UniformFieldSpace(df, [8, 4])
    ufs = F.UniformFieldSpace(df, [8, 8])
    with pytest.raises(F.UniformFieldError):
        ufs.expression_2arg(["nope_x", "nope_y"], "m", lambda a, b: a + b)           # MissingQuantity
