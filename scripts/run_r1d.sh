python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/perf_probe.py cfg3,cfg2,cfg4 exact 2>&1 | tail -8
python bench.py --no-cpu --no-extras | python scripts/pick.py ms_per_step phases_ms roofline e2e.ms_per_step
ncu --set full --clock-control none --import-source on -k regex:k3_gather -s 2 -c 1 -o gpurun_out/prof_k3_r1d -f python scripts/run_once.py cfg3 1 exact 4 > gpurun_out/ncu_k3.log 2>&1
