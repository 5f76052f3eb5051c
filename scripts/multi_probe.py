"""Tuning: per-call wall-clock breakdown of fem2d_galerkin_sample_gep_hcurl_multi on the first N devices: python scripts/multi_probe.py N [calls]"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import fem_2d_b200 as F
import bench
n = int(sys.argv[1]); calls = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = bench.build_product_domain("cfg3"); v = d.view()
glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
nnz = 57557904
h = [torch.empty(nnz, dtype=t).pin_memory() for t in (torch.int32, torch.int32, torch.float64, torch.float64)]
ptrs = (h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), nnz)
for k in range(calls):
    t0 = time.perf_counter()
    F.galerkin_sample_gep_hcurl_multi(v, glq, list(range(n)), out=ptrs)
    ms = 1e3 * (time.perf_counter() - t0)
    tm = (C.c_double * (4 + 4 * n))()
    F._L.fem2d_debug_multi_timing(tm, C.c_uint32(4 + 4 * n))
    print(f"call {k}: {ms:6.2f} ms  host {tm[0]:.2f}", " | ".join(f"sym {tm[4+4*r]:.1f} split {tm[5+4*r]:.1f} num {tm[6+4*r]:.1f}" for r in range(n)), flush=True)
