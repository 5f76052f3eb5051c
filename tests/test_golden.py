"""Frozen golden vectors of the path (tests/golden/gep_*.npz, written by scripts/make_golden.py from the pinned CPU oracle; the
Rust reference cannot run in this image).  CPU: the oracle still reproduces them bit for bit from the stored GLQ nodes.  GPU: the
CUDA path, through the C-ABI, reproduces them bit for bit -- without executing anything under oracle/."""
import glob
import os

import numpy as np
import pytest

import recipes

FILES = sorted(glob.glob(os.path.join(recipes.GOLDEN, "gep_*.npz")))


def _case(path):
    name, shape, basis = os.path.basename(path)[4:-4].rsplit("_", 2)
    z = np.load(path)
    glq = ((z["u_pts"], z["u_w"]), (z["v_pts"], z["v_w"]))
    return name, int(basis[1:]), glq, z


def test_golden_files_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_oracle_reproduces_golden(path):
    import oracle as O
    name, basis, glq, z = _case(path)
    mo, _ = recipes.build_pair(name)
    d = O.Domain.from_mesh(mo)
    g = O.galerkin_sample_gep_hcurl(d, basis=basis, glq=glq)
    assert d.num_dofs == int(z["n_dofs"])
    assert np.array_equal(g.rows, z["rows"]) and np.array_equal(g.cols, z["cols"])
    assert np.array_equal(np.ascontiguousarray(g.a).view(np.uint64), z["a_bits"])
    assert np.array_equal(np.ascontiguousarray(g.b).view(np.uint64), z["b_bits"])


@pytest.mark.gpu
@pytest.mark.parametrize("dedupe", [True, False])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_cuda_path_reproduces_golden(path, dedupe):
    import fem_2d_b200 as F
    name, basis, glq, z = _case(path)
    _, mf = recipes.build_pair(name)
    df = F.Domain.from_mesh(mf)
    plan = F.Plan(df.view(), device=0, dedupe=dedupe)
    rows, cols, a, b = plan.assemble(glq, basis=F.HierPoly if basis == 0 else F.HierMaxOrtho)
    assert df.num_dofs == int(z["n_dofs"])
    assert np.array_equal(rows, z["rows"]) and np.array_equal(cols, z["cols"])
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), z["a_bits"])
    assert np.array_equal(np.ascontiguousarray(b).view(np.uint64), z["b_bits"])
