#!/bin/bash
# sanitizer re-runs (racecheck default PAIR=2, synccheck with immediate barrier count), strict build, refmain on GPU, full GPU suite
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/x16_pytest_gpu.log 2>&1
timeout 240 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "test_feuler_matches_oracle or test_row_synchronisation" > gpurun_out/x16_sanitize_racecheck_default.log 2>&1
echo "rc=$?" >> gpurun_out/x16_sanitize_racecheck_default.log
timeout 240 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "test_feuler_matches_oracle or test_row_synchronisation" > gpurun_out/x16_sanitize_synccheck.log 2>&1
echo "rc=$?" >> gpurun_out/x16_sanitize_synccheck.log
timeout 200 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" > gpurun_out/x16_tune.log 2>&1
echo done > gpurun_out/x16_done.txt
