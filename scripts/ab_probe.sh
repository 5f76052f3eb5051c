# Tuning: A/B of the warp-specialised integrator against k2_exact_kernel for every item (scripts/gpu.sh 900 'bash scripts/ab_probe.sh cfg3,cfg4,hp1m')
W=${1:-cfg3,cfg4,hp1m}
for ws in 0 1; do
  echo "== FEM2D_K2_WS=$ws"
  FEM2D_K2_WS=$ws python scripts/perf_probe.py $W exact 2>&1 | grep -E "workload|ERR|Error" | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['workload'], d['dedupe'], d['items'], d['integrator_ms'], d['scatter_ms'], d['total_ms'])
    except Exception: print(l.strip())"
done
