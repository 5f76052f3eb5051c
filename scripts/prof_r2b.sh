# Round-2 (last session) profile set, trimmed to the kernels that changed (run on a B200: scripts/gpu.sh 1500 'bash scripts/prof_r2b.sh')
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-extras > gpurun_out/r2b_launches.out 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k2_ws -s 2 -c 1 -o gpurun_out/r2b_k2ws_cfg3nd -f python scripts/run_once.py cfg3 0 exact 4 > gpurun_out/ncu_r2b.log 2>&1
$N -k regex:k2_ws -s 2 -c 1 -o gpurun_out/r2b_k2_hp1m -f python scripts/run_once.py hp1m 1 exact 4 >> gpurun_out/ncu_r2b.log 2>&1
$N -k regex:k2_ws -s 2 -c 1 -o gpurun_out/r2b_k2_cfg4 -f python scripts/run_once.py cfg4 1 exact 4 >> gpurun_out/ncu_r2b.log 2>&1
tail -3 gpurun_out/ncu_r2b.log
ls -la gpurun_out/r2b_*.ncu-rep
