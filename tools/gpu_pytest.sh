#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/${TAG:-p1}_pytest_gpu.log 2>&1
echo done > gpurun_out/${TAG:-p1}_done.txt
