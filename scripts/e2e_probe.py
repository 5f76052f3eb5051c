import sys, os, time; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import ctypes as C, numpy as np, torch
import fem_2d_b200 as F, bench
d = bench.build_product_domain("cfg3"); v = d.view()
glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
p0 = F.Plan(v, device=0); n = p0.nnz; del p0
h_rows = torch.empty(n, dtype=torch.int32).pin_memory(); h_cols = torch.empty(n, dtype=torch.int32).pin_memory()
h_a = torch.empty(n, dtype=torch.float64).pin_memory(); h_b = torch.empty(n, dtype=torch.float64).pin_memory()
L = F._L
up, uw, vp, vw = F.Plan._glq_args(glq)
bg = np.array([0], dtype=np.uint64); en = np.array([n], dtype=np.uint64)
P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
for it in range(4):
    t0 = time.perf_counter()
    plan = F.Plan(v, device=0)
    t1 = time.perf_counter()
    st = L.fem2d_assemble_ranges(plan._h, 0, 0, 1, 0, P(up, C.c_double), P(uw, C.c_double), C.c_uint32(8), P(vp, C.c_double), P(vw, C.c_double), C.c_uint32(8),
                                 C.c_uint32(1), P(bg, C.c_uint64), P(en, C.c_uint64), C.c_void_p(h_rows.data_ptr()), C.c_void_p(h_cols.data_ptr()),
                                 C.c_void_p(h_a.data_ptr()), C.c_void_p(h_b.data_ptr()))
    t2 = time.perf_counter()
    st = L.fem2d_assemble_ranges(plan._h, 0, 0, 1, 0, P(up, C.c_double), P(uw, C.c_double), C.c_uint32(8), P(vp, C.c_double), P(vw, C.c_double), C.c_uint32(8),
                                 C.c_uint32(1), P(bg, C.c_uint64), P(en, C.c_uint64), C.c_void_p(h_rows.data_ptr()), C.c_void_p(h_cols.data_ptr()),
                                 C.c_void_p(h_a.data_ptr()), C.c_void_p(h_b.data_ptr()))
    t3 = time.perf_counter()
    st = L.fem2d_assemble_ranges(plan._h, 0, 0, 1, 0, P(up, C.c_double), P(uw, C.c_double), C.c_uint32(8), P(vp, C.c_double), P(vw, C.c_double), C.c_uint32(8),
                                 C.c_uint32(1), P(bg, C.c_uint64), P(en, C.c_uint64), None, C.c_void_p(h_cols.data_ptr()),
                                 C.c_void_p(h_a.data_ptr()), C.c_void_p(h_b.data_ptr()))
    t3b = time.perf_counter()
    print(f"   without rows: {1e3*(t3b-t3):.2f}")
    t3 = time.perf_counter()
    del plan
    t4 = time.perf_counter()
    print(f"symbolic {1e3*(t1-t0):.2f}  first numeric+d2h {1e3*(t2-t1):.2f}  second {1e3*(t3-t2):.2f}  free {1e3*(t4-t3):.2f}")
