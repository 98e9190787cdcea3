#!/bin/bash
# same-box A/B with the fast (default-kind-only) build: default kernel vs the full-width-tile variant v2
# (EULERB200_XC=1: column double-buffered, top warp one plane ahead)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
export EULERB200_LIB=$PWD/sundials-manyvector-demo_b200/libeulerb200_fast.so
timeout 300 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "XC=1" "" "XC=1" "XC=1 PAIR=1" "XC=1 PAIR=0" "XC=1 CTAS=8800" "XC=1 CTAS=4400" > gpurun_out/x25_tune.log 2>&1
timeout 100 python tools/tune2.py --n 512 512 512 --nchem 0 --steps 5 --env "" "XC=1" "VARIANT=1" "XC=1 VARIANT=1" > gpurun_out/x25_tune_nchem0.log 2>&1
EULERB200_XC=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "test_feuler_matches_oracle or test_row_sync" > gpurun_out/x25_pytest_xc.log 2>&1
EULERB200_XC=1 timeout 400 compute-sanitizer --tool racecheck --target-processes all --print-limit 20 \
     python -m pytest tests/test_gpu_parity.py -x -q -k "test_feuler_matches_oracle" > gpurun_out/x25_sanitize_racecheck_xc.log 2>&1
echo "rc=$?" >> gpurun_out/x25_sanitize_racecheck_xc.log
echo done > gpurun_out/x25_done.txt
