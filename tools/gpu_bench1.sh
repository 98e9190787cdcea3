#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 600 python bench.py > gpurun_out/x12_bench_n1.json 2> gpurun_out/x12_bench_n1.err
timeout 300 python bench.py --ic problem --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/x12_bench_n1_blastic.json 2> gpurun_out/x12_bench_n1_blastic.err
timeout 300 python bench.py --workload rayleigh_taylor --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/x12_bench_rt.json 2> gpurun_out/x12_bench_rt.err
timeout 300 python bench.py --workload hurricane_yz --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/x12_bench_hurricane.json 2> gpurun_out/x12_bench_hurricane.err
timeout 300 python bench.py --workload linear_advection_x --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/x12_bench_advection.json 2> gpurun_out/x12_bench_advection.err
echo done > gpurun_out/x12_done.txt
