#!/bin/bash
# final one-GPU evidence on the build with the leaner arithmetic (r2p): GPU test suite, default bench line, full ncu
# capture of the hot kernel, ncu launch list of the bench command, the other BASELINE.json grid shapes, blast
# initial condition, reference arm, smoke -- most important first (the GPU budget of the round is nearly spent)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-f4}
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1
timeout 400 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
NCU_TAG=${T}_fused bash tools/gpu_ncu_fused.sh
rm -f gpurun_out/${T}_fused.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_bench_512cube.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${T}_ncu_list.log 2>&1
for wl in rayleigh_taylor hurricane_yz linear_advection_x; do
  timeout 120 python bench.py --workload $wl --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench_$wl.json 2> gpurun_out/${T}_bench_$wl.err
done
timeout 120 python bench.py --ic problem --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench_n1_blast_ic.json 2> gpurun_out/${T}_bench_n1_blast_ic.err
timeout 150 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo done > gpurun_out/${T}_done.txt
