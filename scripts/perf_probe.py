"""GPU probe used while tuning: per-phase timings of several workloads (dedupe on/off)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import fem_2d_b200 as F
import bench

modes = {"exact": F.MODE_EXACT, "sumfact": F.MODE_SUMFACT, "dmma": F.MODE_DMMA}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg3", "cfg2", "cfg4"]
mlist = sys.argv[2].split(",") if len(sys.argv) > 2 else ["exact"]
for wl in which:
    d = bench.build_product_domain(wl)
    v = d.view()
    g = bench.WORKLOADS[wl]["glq"]
    glq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
    for dedupe in (1, 0):
        plan = F.Plan(v, device=0, dedupe=bool(dedupe)); plan.set_phase_timing(True)
        da = torch.empty(plan.nnz, dtype=torch.float64, device="cuda"); db = torch.empty_like(da)
        for mname in mlist:
            try:
                for _ in range(3):
                    plan.assemble_device(glq, da.data_ptr(), db.data_ptr(), mode=modes[mname])
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    plan.assemble_device(glq, da.data_ptr(), db.data_ptr(), mode=modes[mname])
                    ts.append(plan.last_timing())
                t = {k: float(np.median([x[k] for x in ts])) for k in ("sampler_ms", "integrator_ms", "scatter_ms", "total_ms")}
                print(json.dumps({"workload": wl, "mode": mname, "dedupe": dedupe, "dofs": plan.n_dofs, "nnz": plan.nnz, "pairs": plan.info["n_pairs"],
                                  "classes": plan.info["n_classes"], "items": plan.info["n_work_items"], **{k: round(x, 4) for k, x in t.items()},
                                  "Gnnz_s": round(2 * plan.nnz / t["total_ms"] / 1e6, 2)}), flush=True)
            except F.BackendError as e:
                print(wl, mname, dedupe, "ERR", e)
        del plan, da, db
