# Tuning: per-phase timings of the product library and of every tuning build under fem_2d_b200/_variants/ (scripts/gpu.sh 900 'bash scripts/variants_probe.sh cfg3,cfg4,hp1m')
W=${1:-cfg3,cfg4,hp1m}
echo "== product"; python scripts/perf_probe.py $W exact 2>&1 | grep workload | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'], d['dedupe'], d['items'], d['integrator_ms'], d['scatter_ms'], d['total_ms'])"
for v in fem_2d_b200/_variants/*/; do
  n=$(basename $v); echo "== $n"
  FEM2D_LIB=$v/libfem2d_b200.so python scripts/perf_probe.py $W exact 2>&1 | grep -E "workload|ERR|Error" | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['workload'], d['dedupe'], d['items'], d['integrator_ms'], d['scatter_ms'], d['total_ms'])
    except Exception: print(l.strip())"
done
