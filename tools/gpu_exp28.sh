#!/bin/bash
# same-box A/B: AG instantiation (boundary tiles read the per-cell arrays) for every launch vs only boundary-heavy ones
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "AG_FRAC=0" "" "AG_FRAC=0" "CTAS=8800" "AG_FRAC=0 CTAS=8800" > gpurun_out/x28_tune.log 2>&1
timeout 100 python tools/tune2.py --n 512 512 512 --nchem 0 --steps 5 --env "" "AG_FRAC=0" > gpurun_out/x28_tune_nchem0.log 2>&1
timeout 100 python tools/tune2.py --n 256 256 256 --nchem 0 --steps 10 --env "" "AG_FRAC=0" > gpurun_out/x28_tune_256.log 2>&1
echo done > gpurun_out/x28_done.txt
