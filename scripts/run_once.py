"""Runs a few numeric steps of one workload (for ncu captures): python scripts/run_once.py <workload> <dedupe> <mode> [steps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import fem_2d_b200 as F
import bench
wl, dedupe, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
modes = {"exact": F.MODE_EXACT, "sumfact": F.MODE_SUMFACT, "dmma": F.MODE_DMMA}
d = bench.build_product_domain(wl); v = d.view()
g = bench.WORKLOADS[wl]["glq"]
glq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
plan = F.Plan(v, device=0, dedupe=bool(dedupe)); plan.set_phase_timing(True)
a = torch.empty(plan.nnz, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
for _ in range(steps):
    plan.assemble_device(glq, a.data_ptr(), b.data_ptr(), mode=modes[mode])
torch.cuda.synchronize()
print(plan.last_timing())
