# Round-end check on a B200: GPU tests, smoke, the default bench (scripts/gpu.sh 1500 'bash scripts/run_final.sh')
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
python scripts/pick.py ms_per_step value roofline.frac roofline_fp64.frac roofline_fp64.kernel_ms e2e.ms_per_step gpu_launches < gpurun_out/r2b_bench_n1.json
