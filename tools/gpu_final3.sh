#!/bin/bash
# final one-GPU evidence of the round: GPU test suite, the default bench line (both initial conditions), the
# other BASELINE.json grid shapes, the reference arm
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/f3_pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/f3_bench_n1.json 2> gpurun_out/f3_bench_n1.err
timeout 200 python bench.py --ic problem --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/f3_bench_n1_blast_ic.json 2> gpurun_out/f3_bench_n1_blast_ic.err
for wl in rayleigh_taylor hurricane_yz linear_advection_x; do
  timeout 200 python bench.py --workload $wl --no-cpu-baseline --steps 5 > gpurun_out/f3_bench_$wl.json 2> gpurun_out/f3_bench_$wl.err
done
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/f3_bench_reference_arm.json 2> gpurun_out/f3_bench_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1
echo done > gpurun_out/f3_done.txt
