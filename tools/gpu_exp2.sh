#!/bin/bash
# Round-2 experiment 2: lean WENO (default now) and the fluid / species split at several CTA sizes.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "SPLIT=1" "SPLIT=1 VARIANT_T=2" "SPLIT=1 VARIANT_T=3" \
   "SPLIT=1 VARIANT_F=2 VARIANT_T=2" "SPLIT=1 VARIANT_F=2 VARIANT_T=3" "VARIANT=2" "SPLIT=1 VARIANT_T=3 PAIR=0" > gpurun_out/x2_tune.log 2>&1
timeout 100 python tools/tune2.py --n 512 512 512 --nchem 0 --steps 5 --env "" "VARIANT=2" > gpurun_out/x2_tune_nchem0.log 2>&1
# per-launch times of the split (fluid launch vs species launch)
EULERB200_SPLIT=1 EULERB200_VARIANT_T=3 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rhs_fused --csv \
   --log-file gpurun_out/x2_launches_split_t640.csv python tools/tune2.py --n 512 512 512 --nchem 10 --steps 1 --env "SPLIT=1 VARIANT_T=3" "SPLIT=1 VARIANT_T=2" > gpurun_out/x2_ncu.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/x2_pytest_gpu.log 2>&1
EULERB200_SPLIT=1 EULERB200_VARIANT_T=3 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/x2_pytest_gpu_split.log 2>&1
echo done > gpurun_out/x2_done.txt
