"""Writes tests/golden/gep_<recipe>.npz: the A/B matrices of small recipes as assembled by the CPU oracle (the restatement of
galerkin_sample_gep_hcurl pinned on the reference's fixtures, tests/test_oracle_pinned.py).  The Rust reference itself cannot
run in this image (no rustc/cargo), so these frozen oracle outputs are the golden vectors of the path: values are stored as raw
IEEE-754 bit patterns (uint64) because parity is bit-exact.   python scripts/make_golden.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
import recipes

CASES = [("readme", 8, 8, 0), ("slepc", 8, 8, 0), ("cfg4_small", 5, 9, 0), ("slepc", 12, 12, 1)]   # (recipe, nu, nv, basis)
for name, nu, nv, basis in CASES:
    mo, _ = recipes.build_pair(name)
    d = O.Domain.from_mesh(mo)
    glq = (O.gauss_quadrature_points(nu), O.gauss_quadrature_points(nv))
    g = O.galerkin_sample_gep_hcurl(d, basis=basis, glq=glq)
    out = os.path.join(recipes.GOLDEN, f"gep_{name}_{nu}x{nv}_b{basis}.npz")
    np.savez_compressed(out, rows=g.rows.astype(np.uint32), cols=g.cols.astype(np.uint32), a_bits=np.ascontiguousarray(g.a).view(np.uint64),
                        b_bits=np.ascontiguousarray(g.b).view(np.uint64), n_dofs=np.int64(d.num_dofs),
                        u_pts=glq[0][0], u_w=glq[0][1], v_pts=glq[1][0], v_w=glq[1][1])
    print(out, d.num_dofs, len(g.rows), os.path.getsize(out))
