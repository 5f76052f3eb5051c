"""Tuning: cycle breakdown of the warp-specialised integrator (needs a -DFEM2D_WS_PROFILE build): python scripts/ws_profile.py <workload> <dedupe>"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import fem_2d_b200 as F
import bench
wl, dedupe = sys.argv[1], int(sys.argv[2])
d = bench.build_product_domain(wl); v = d.view()
g = bench.WORKLOADS[wl]["glq"]
glq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
plan = F.Plan(v, device=0, dedupe=bool(dedupe)); plan.set_phase_timing(True)
a = torch.empty(plan.nnz, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
for _ in range(2):
    plan.assemble_device(glq, a.data_ptr(), b.data_ptr())
torch.cuda.synchronize()
out = (C.c_uint64 * 16)()
F._L.fem2d_debug_ws_profile(out, 1)
plan.assemble_device(glq, a.data_ptr(), b.data_ptr())
torch.cuda.synchronize()
F._L.fem2d_debug_ws_profile(out, 1)
t = plan.last_timing()
names = ["prod_setup", "prod_wait_empty", "prod_stage", "cons_wait_full", "cons_contract", "items", "chunks", "cons_wait_first",
         "setup_claim", "setup_desc", "setup_offsets_tables", "setup_cols", "setup_orders", "-", "-", "-"]
vals = [int(x) for x in out]
ctas = 296
print(wl, dedupe, "integrator_ms", round(t["integrator_ms"], 4), "kernel cycles", int(t["integrator_ms"] * 1.965e6), "cycles per CTA:", {n: vals[k] // ctas for k, n in enumerate(names)})
