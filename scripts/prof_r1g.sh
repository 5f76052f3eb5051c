# Round-1 final profile set (run on a B200: scripts/gpu.sh 1500 'bash scripts/prof_r1g.sh')
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-extras > gpurun_out/launches_r1g.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_gather -s 2 -c 1 -o gpurun_out/prof_k3_r1g -f python scripts/run_once.py cfg3 1 exact 4 > gpurun_out/ncu_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_nd_r1g -f python scripts/run_once.py cfg3 0 exact 4 >> gpurun_out/ncu_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 4 -c 2 -o gpurun_out/prof_k2_cfg4_r1g -f python scripts/run_once.py cfg4 1 exact 4 >> gpurun_out/ncu_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_exact -s 2 -c 1 -o gpurun_out/prof_k2_dd_r1g -f python scripts/run_once.py cfg3 1 exact 4 >> gpurun_out/ncu_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_gather -s 2 -c 1 -o gpurun_out/prof_k3_nd_r1g -f python scripts/run_once.py cfg3 0 exact 4 >> gpurun_out/ncu_r1g.log 2>&1
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err
tail -c 400 gpurun_out/bench_r1g.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1g.json 2> gpurun_out/bench_ref_r1g.err
tail -c 600 gpurun_out/bench_ref_r1g.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
