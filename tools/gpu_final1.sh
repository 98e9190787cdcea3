#!/bin/bash
# Round-2 evidence run on one B200: bench lines (default workload with both initial conditions, the other
# BASELINE.json grid shapes), the reference arm, the ncu launch list of the bench command, one full capture
# of the hot kernel, the GPU test suite.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/f1_pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/f1_bench_n1.json 2> gpurun_out/f1_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/f1_bench_reference_arm.json 2> gpurun_out/f1_bench_reference_arm.err
timeout 300 python bench.py --ic problem --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/f1_bench_n1_blast_ic.json 2> gpurun_out/f1_bench_n1_blast_ic.err
for wl in rayleigh_taylor hurricane_yz linear_advection_x; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 5 > gpurun_out/f1_bench_$wl.json 2> gpurun_out/f1_bench_$wl.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f1_launches_bench_512cube.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/f1_ncu_list.log 2>&1
NCU_TAG=f1_fused bash tools/gpu_ncu_fused.sh
rm -f gpurun_out/f1_fused.ncu-rep
echo done > gpurun_out/f1_done.txt
