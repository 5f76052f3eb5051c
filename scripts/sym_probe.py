import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import fem_2d_b200 as F, bench, time
d = bench.build_product_domain(sys.argv[1] if len(sys.argv) > 1 else "cfg3"); v = d.view()
for i in range(4):
    t0=time.perf_counter(); p = F.Plan(v, device=0); t1=time.perf_counter()
    print("plan", round((t1-t0)*1e3,2), "host", p.info["symbolic_host_us"]/1e3, "dev", p.info["symbolic_device_us"]/1e3, "nnz", p.nnz); del p
