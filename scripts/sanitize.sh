# compute-sanitizer passes over the GPU tests (run on a B200: scripts/gpu.sh 1700 'bash scripts/sanitize.sh'): memcheck over the parity, fuzz, golden, field,
# PETSc and multi-device tests; racecheck and initcheck over the golden tests and one hp-mesh parity case (persistent integrator: mbarrier ring, packs)
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_golden.py tests/test_fields_gpu.py tests/test_petsc_gpu.py -m gpu -x -q -k "not cfg4_full" 2>&1 | tail -4
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_multi_device_gpu.py -m gpu -x -q -k "small or error" 2>&1 | tail -4
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_golden.py "tests/test_gpu_parity.py::test_exact_bit_identical" -m gpu -x -q -k "golden or cfg4_small or edge_order" 2>&1 | tail -4
compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_golden.py "tests/test_gpu_parity.py::test_exact_bit_identical" -m gpu -x -q -k "golden or cfg4_small or edge_order" 2>&1 | tail -4
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest "tests/test_gpu_parity.py::test_exact_bit_identical" -m gpu -x -q -k "cfg4_small or slepc" 2>&1 | tail -4
# the persistent warp-specialised integrator (mbarrier ring, packs, one and two staging warps): throughput-shape fuzz mesh (14 611 DoFs, tile_p = 4)
for tool in racecheck initcheck synccheck; do
compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_fuzz_gpu.py -m gpu -x -q -k "throughput" 2>&1 | tail -3
FEM2D_K2_WS_PROD=1 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_fuzz_gpu.py -m gpu -x -q -k "throughput and 0" 2>&1 | tail -3
done
