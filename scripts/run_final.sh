# Round-end check on a B200: GPU tests, smoke, a short bench (scripts/gpu.sh 1200 'bash scripts/run_final.sh')
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python scripts/perf_probe.py cfg3,cfg2,cfg4 exact 2>&1 | grep workload | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'], d['dedupe'], d['items'], d['integrator_ms'], d['total_ms'])"
python bench.py --no-cpu --e2e-steps 2 2>/dev/null | python scripts/pick.py ms_per_step value roofline.frac e2e.ms_per_step gpu_launches other_configs
