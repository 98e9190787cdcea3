#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python -m pytest tests/test_gpu_strict.py -x -q > gpurun_out/x18_pytest_strict.log 2>&1
for pair in 0 2; do
EULERB200_PAIR=$pair timeout 200 compute-sanitizer --tool synccheck --print-limit 3000 python -m pytest tests/test_gpu_parity.py -x -q -k "test_illegal_state" > gpurun_out/x18_synccheck_pair$pair.log 2>&1
echo "rc=$?" >> gpurun_out/x18_synccheck_pair$pair.log
done
echo done > gpurun_out/x18_done.txt
