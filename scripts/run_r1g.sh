# Integrator tile-shape A/B (run on a B200: scripts/gpu.sh 1200 'bash scripts/run_r1g.sh')
set -x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python scripts/perf_probe.py cfg3,cfg2,cfg4 exact 2>&1 | tail -8 | tee gpurun_out/probe_r1g.jsonl
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_nd_r1g -f python scripts/run_once.py cfg3 0 exact 4 > gpurun_out/ncu_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 4 -c 2 -o gpurun_out/prof_k2_cfg4_r1g -f python scripts/run_once.py cfg4 1 exact 4 >> gpurun_out/ncu_r1g.log 2>&1
tail -3 gpurun_out/ncu_r1g.log
