#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "CTAS=7990" "CTAS=3996" "PAIR=0" "KERNEL=0" > gpurun_out/x21_tune.log 2>&1
timeout 100 python tools/tune2.py --n 512 512 512 --nchem 0 --steps 5 --env "" "VARIANT=2" > gpurun_out/x21_tune_nchem0.log 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/x21_pytest_gpu.log 2>&1
echo done > gpurun_out/x21_done.txt
