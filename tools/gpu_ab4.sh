#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
timeout 70 ncu --set full --clock-control none --import-source on -k regex:tracer_kernel -s 2 -c 1 -f -o gpurun_out/r1i_tracer \
   python tools/tune.py --n 256 256 256 --nchem 10 --variants 1 --split 2 --steps 1 > gpurun_out/ab4.log 2>&1
echo done > gpurun_out/ab4_done.txt
