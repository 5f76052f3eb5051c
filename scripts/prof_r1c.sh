set -x
python scripts/perf_probe.py cfg3,cfg2,cfg4 exact,sumfact,dmma > gpurun_out/probe_r1c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-extras > gpurun_out/launches_r1c.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_gather -s 2 -c 1 -o gpurun_out/prof_k3_r1c -f python scripts/run_once.py cfg3 1 exact 4 > gpurun_out/ncu_k3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_r1c -f python scripts/run_once.py cfg3 0 exact 4 > gpurun_out/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_cfg4_r1c -f python scripts/run_once.py cfg4 1 exact 4 > gpurun_out/ncu_k2c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_dd_r1c -f python scripts/run_once.py cfg3 1 exact 4 > gpurun_out/ncu_k2dd.log 2>&1
cat gpurun_out/probe_r1c.log
