# Tuning: power-of-two scales folded into the weights or not (scripts/gpu.sh 900 'bash scripts/fold_probe.sh')
W=${1:-cfg3,cfg4,hp1m}
for fold in 0 3; do
echo "== FEM2D_K2_WS_FOLD=$fold"
FEM2D_K2_WS_FOLD=$fold python scripts/perf_probe.py $W exact 2>&1 | grep -E "workload|ERR|Error" | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['workload'], d['dedupe'], d['items'], d['integrator_ms'], d['scatter_ms'], d['total_ms'])
    except Exception: print(l.strip())"
done
python -m pytest tests/test_fuzz_gpu.py tests/test_golden.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
