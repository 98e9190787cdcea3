#!/bin/bash
# Second one-shot GPU session: prefetch variants A/B, then bench + parity tests under the fastest one.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 100 python tools/tune.py --n 512 512 512 --nchem 10 --variants 1 5 6 7 8 9 10 4 1 --pair 2 --steps 5 > gpurun_out/ab2.log 2>&1
best=$(grep "nchem=10" gpurun_out/ab2.log | sort -k7 -n | head -1 | sed 's/.*variant=\([0-9]*\).*/\1/')
echo "best=$best" > gpurun_out/ab2_best.txt
EULERB200_VARIANT=$best timeout 70 python bench.py > gpurun_out/bench_best.json 2> gpurun_out/bench_best.err
timeout 30 python tools/tune.py --n 512 512 512 --nchem 0 --variants 1 6 7 1 --pair 2 --steps 5 > gpurun_out/ab2_nvar5.log 2>&1
EULERB200_VARIANT=$best timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_and_halo.py -x -q > gpurun_out/pytest_best.log 2>&1
echo done > gpurun_out/ab2_done.txt
