#!/bin/bash
# evidence run on one B200 for the current default kernel: GPU test suite, the default bench line, the ncu launch
# list of the bench command, one full capture of the hot kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-f2}
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_bench_512cube.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${T}_ncu_list.log 2>&1
NCU_TAG=${T}_fused bash tools/gpu_ncu_fused.sh
rm -f gpurun_out/${T}_fused.ncu-rep
echo done > gpurun_out/${T}_done.txt
