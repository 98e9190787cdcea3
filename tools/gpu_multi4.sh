#!/bin/bash
# 2-GPU bench line on the final build (weak scaling, rank-seam parity), short form: no e2e / cpu_baseline legs
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/m4_bench_n2.json 2> gpurun_out/m4_bench_n2.err
echo done > gpurun_out/m4_done.txt
