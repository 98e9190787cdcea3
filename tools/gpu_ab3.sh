#!/bin/bash
# Third one-shot GPU session (what is left of the budget): split mode (tracer_kernel) against the fused kernel.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
timeout 100 python tools/tune.py --n 512 512 512 --nchem 10 --variants 1 --split 0 2 --steps 4 > gpurun_out/ab3.log 2>&1
echo done > gpurun_out/ab3_done.txt
