#!/bin/bash
# same-box A/B: default kernel vs the bulk-copy staging variant (and CTA-wide barriers alone, which STAGE implies)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "STAGE=1" "PAIR=0" "" "STAGE=1" > gpurun_out/x22_tune.log 2>&1
EULERB200_STAGE=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/x22_pytest_stage.log 2>&1
NCU_TAG=x22_stage NCU_ENV="STAGE=1" bash tools/gpu_ncu_fused.sh
rm -f gpurun_out/x22_stage.ncu-rep gpurun_out/x22_stage.source.csv
echo done > gpurun_out/x22_done.txt
