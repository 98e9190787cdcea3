#!/bin/bash
# Round-2 experiment 10: trimmed species loops (branch-free, running pointers, compiled-in tile shape), one-Newton WENO reciprocal.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/x10_pytest_gpu.log 2>&1
timeout 300 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "CHEMT=1" "PAIR=1" "VARIANT=2" "SPLIT=1" > gpurun_out/x10_tune.log 2>&1
EULERB200_LIB=$PWD/sundials-manyvector-demo_b200/libeulerb200_n1.so timeout 200 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" >> gpurun_out/x10_tune_n1.log 2>&1
timeout 100 python tools/tune2.py --n 512 512 512 --nchem 0 --steps 5 --env "" "VARIANT=2" > gpurun_out/x10_tune_nchem0.log 2>&1
timeout 200 python tools/configs_bench.py > gpurun_out/x10_configs.log 2>&1
echo done > gpurun_out/x10_done.txt
