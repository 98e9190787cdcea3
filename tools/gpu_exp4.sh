#!/bin/bash
# Round-2 experiment 4: pair-interleaved species copy (chemT) on / off, fused and split.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 400 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "NO_CHEMT=1" "SPLIT=1" "SPLIT=1 VARIANT_T=2" "SPLIT=1 VARIANT_T=3" \
   "VARIANT=2" "SPLIT=1 VARIANT_F=2 VARIANT_T=2" "PAIR=1" "SPLIT=1 VARIANT_T=2 PAIR=1" > gpurun_out/x4_tune.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/x4_launches.csv python tools/tune2.py --n 512 512 512 --nchem 10 --steps 1 --env "" "SPLIT=1 VARIANT_T=2" "SPLIT=1 VARIANT_T=3" > gpurun_out/x4_ncu.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/x4_pytest_gpu.log 2>&1
echo done > gpurun_out/x4_done.txt
