"""fem2d_galerkin_sample_gep_hcurl_multi: ONE GEP out of a multi-device call of one process (galerkin.rs:33-40 returns one GEP; linalg.rs:59-81
merges everything into it).  The arrays must equal the single-device result bit for bit for any device count.  On a box with one GPU the
device list repeats device 0, which still exercises the row-block partition, the restricted integrator of every block and the placement of
every block's slices at their slot positions; with more GPUs visible the first min(n, 4) distinct devices are used as well."""
import numpy as np
import pytest

import recipes

pytestmark = pytest.mark.gpu

import fem_2d_b200 as F  # noqa: E402
import oracle as O  # noqa: E402


def _single(df, glq, **kw):
    plan = F.Plan(df.view(), device=0)
    return plan.assemble(glq, **kw)


@pytest.mark.parametrize("name,g", [("readme", 8), ("slepc", 8), ("cfg4_small", 12)])
def test_multi_equals_single_and_oracle_small(name, g):
    mo, mf = recipes.build_pair(name)
    do, df = O.Domain.from_mesh(mo), F.Domain.from_mesh(mf)
    glq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
    ref = O.galerkin_sample_gep_hcurl(do, glq=glq)
    for devs in ([0], [0, 0], [0, 0, 0]):
        rows, cols, a, b = F.galerkin_sample_gep_hcurl_multi(df, glq, devs)
        assert np.array_equal(rows, ref.rows) and np.array_equal(cols, ref.cols)
        assert np.array_equal(a.view(np.uint64), ref.a.view(np.uint64)) and np.array_equal(b.view(np.uint64), ref.b.view(np.uint64))


def test_multi_on_the_hp_mesh_and_second_basis():
    """BASELINE configs[3] (restricted work items with several tile ranges, local-desc blocks) on 3 blocks, both basis spaces."""
    df = F.Domain.from_mesh(recipes.mesh_cfg4(recipes.api("product")))
    glq = (F.gauss_quadrature_points(12), F.gauss_quadrature_points(12))
    for basis in (F.HierPoly, F.HierMaxOrtho):
        r1, c1, a1, b1 = _single(df, glq, basis=basis)
        rows, cols, a, b = F.galerkin_sample_gep_hcurl_multi(df, glq, [0, 0, 0], basis=basis)
        assert np.array_equal(rows, r1) and np.array_equal(cols, c1)
        assert np.array_equal(a.view(np.uint64), a1.view(np.uint64)) and np.array_equal(b.view(np.uint64), b1.view(np.uint64))


def test_multi_full_size_all_visible_devices():
    """BASELINE configs[2] (1.18 M DoFs) from one process on every visible GPU (up to 4; device 0 twice on a single-GPU box)."""
    n = min(F.device_count(), 4)
    devs = list(range(n)) if n > 1 else [0, 0]
    df = F.Domain.from_mesh(recipes.mesh_cfg3(recipes.api("product")))
    glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
    r1, c1, a1, b1 = _single(df, glq)
    rows, cols, a, b = F.galerkin_sample_gep_hcurl_multi(df, glq, devs)
    assert len(rows) == 57557904
    assert np.array_equal(rows, r1) and np.array_equal(cols, c1)
    assert np.array_equal(a.view(np.uint64), a1.view(np.uint64)) and np.array_equal(b.view(np.uint64), b1.view(np.uint64))


def test_multi_error_statuses():
    df = F.Domain.from_mesh(recipes.RECIPES["readme"](recipes.api("product")))
    glq3 = (F.gauss_quadrature_points(4)[0][:3], F.gauss_quadrature_points(4)[1][:3])
    with pytest.raises(F.GalerkinSamplingError) as e:
        F.galerkin_sample_gep_hcurl_multi(df, (glq3, glq3), [0])
    assert e.value.kind == F.GalerkinSamplingError.InvalidGLQSettings
    with pytest.raises(F.BackendError):
        F.galerkin_sample_gep_hcurl_multi(df, (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8)), [99])
