#!/bin/bash
# 2-GPU verification of the final build: the multi-GPU test file
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/m3_pytest_multi_n2.log 2>&1
echo done > gpurun_out/m3_done.txt
