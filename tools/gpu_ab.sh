#!/bin/bash
# One-shot GPU session: A/B of the kernel variants, ncu capture of the default kernel, launch list, bench line.
# Everything is bounded by `timeout`; results land in gpurun_out/ as they are produced.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab_gpu.txt 2>&1
timeout 200 python tools/tune.py --n 512 512 512 --nchem 10 --variants 1 4 6 7 5 1 --pair 2 --steps 5 > gpurun_out/ab.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:rhs_fused -s 3 -c 1 -f -o gpurun_out/r1g_default \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1g_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo done > gpurun_out/ab_done.txt
