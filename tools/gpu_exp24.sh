#!/bin/bash
# same-box A/B: L2 prefetch of the plane ahead by the tile's idle top warp (EULERB200_PF=D)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "PF=1" "PF=2" "PF=4" "" "PF=1" > gpurun_out/x24_tune.log 2>&1
timeout 100 python tools/tune2.py --n 512 512 512 --nchem 0 --steps 5 --env "" "PF=1" "PF=2" > gpurun_out/x24_tune_nchem0.log 2>&1
EULERB200_PF=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/x24_pytest_pf.log 2>&1
echo done > gpurun_out/x24_done.txt
