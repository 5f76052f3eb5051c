# Round-1 profile set (run on a B200: scripts/gpu.sh 1500 'bash scripts/prof_r1f.sh')
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-extras > gpurun_out/launches_r1f.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_gather -s 2 -c 1 -o gpurun_out/prof_k3_r1f -f python scripts/run_once.py cfg3 1 exact 4 > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_nd_r1f -f python scripts/run_once.py cfg3 0 exact 4 >> gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 4 -c 2 -o gpurun_out/prof_k2_cfg4_r1f -f python scripts/run_once.py cfg4 1 exact 4 >> gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_dd_r1f -f python scripts/run_once.py cfg3 1 exact 4 >> gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_gather -s 2 -c 1 -o gpurun_out/prof_k3_nd_r1f -f python scripts/run_once.py cfg3 0 exact 4 >> gpurun_out/ncu.log 2>&1
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
tail -c 600 gpurun_out/bench_r1f.json
