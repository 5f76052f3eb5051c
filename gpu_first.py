# scratch: first GPU run -- timing of cfg 3 and FP64 peaks
import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import fem_2d_b200 as F
import recipes
print(F.version(), 'devices', F.device_count())
print('fp64 dfma GF/s', F.fp64_peak(0, 0), 'dmul+dadd GF/s', F.fp64_peak(0, 1))
api = recipes.api('product')
for levels in (4, 6):
    t = time.time(); m = recipes.mesh_cfg3(api, levels=levels); d = F.Domain.from_mesh(m); v = d.view(); t1 = time.time()
    print('levels', levels, 'elems', m.num_elems, 'dofs', d.num_dofs, 'host build s', t1 - t)
    glq = (F.gauss_quadrature_points(8), F.gauss_quadrature_points(8))
    for dedupe in (True, False):
        t = time.time(); plan = F.Plan(v, device=0, dedupe=dedupe); t2 = time.time()
        print(' dedupe', dedupe, 'symbolic s', t2 - t, plan.info)
        import torch
        da = torch.empty(plan.nnz, dtype=torch.float64, device='cuda'); db = torch.empty_like(da)
        for it in range(4):
            plan.assemble_device(glq, da.data_ptr(), db.data_ptr())
            print('   ', plan.last_timing())
        print('   checksum', float(da.sum()), float(db.sum()))
        del plan
