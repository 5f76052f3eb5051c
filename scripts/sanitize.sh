#!/bin/bash
# compute-sanitizer passes over the small GPU tests (run on a GPU box: scripts/gpu.sh 2400 'bash scripts/sanitize.sh')
set -x
compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_fields_gpu.py -x -q -m gpu 2>&1 | tail -6
compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_golden.py -x -q -m gpu -k "readme or cfg4" 2>&1 | tail -6
compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 8 python -m pytest tests/test_golden.py -x -q -m gpu -k "readme or cfg4 or slepc_12" 2>&1 | tail -6
