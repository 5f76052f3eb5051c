# Round-2 profile set (run on a B200: scripts/gpu.sh 1700 'bash scripts/prof_r2.sh')
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-extras > gpurun_out/r2_launches.out 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k3_gather -s 2 -c 1 -o gpurun_out/r2_k3_cfg3 -f python scripts/run_once.py cfg3 1 exact 4 > gpurun_out/ncu_r2.log 2>&1
$N -k regex:k2_ws -s 2 -c 1 -o gpurun_out/r2_k2ws_cfg3nd -f python scripts/run_once.py cfg3 0 exact 4 >> gpurun_out/ncu_r2.log 2>&1
$N -k regex:k2_ws -s 2 -c 1 -o gpurun_out/r2_k2_hp1m -f python scripts/run_once.py hp1m 1 exact 4 >> gpurun_out/ncu_r2.log 2>&1
$N -k regex:k3_gather -s 2 -c 1 -o gpurun_out/r2_k3_hp1m -f python scripts/run_once.py hp1m 1 exact 4 >> gpurun_out/ncu_r2.log 2>&1
$N -k regex:k2_ws -s 2 -c 1 -o gpurun_out/r2_k2_cfg4 -f python scripts/run_once.py cfg4 1 exact 4 >> gpurun_out/ncu_r2.log 2>&1
$N -k regex:k2_exact -s 2 -c 1 -o gpurun_out/r2_k2_cfg3dd -f python scripts/run_once.py cfg3 1 exact 4 >> gpurun_out/ncu_r2.log 2>&1
$N -k "regex:sumfact|gram" -s 4 -c 3 -o gpurun_out/r2_sumfact_cfg3nd -f python scripts/run_once.py cfg3 0 sumfact 4 >> gpurun_out/ncu_r2.log 2>&1
$N -k "regex:dmma|gram" -s 4 -c 3 -o gpurun_out/r2_dmma_cfg3nd -f python scripts/run_once.py cfg3 0 dmma 4 >> gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
ls -la gpurun_out/*.ncu-rep
