#!/bin/bash
# 2-GPU call: multi-GPU parity tests, bench at N=2 (seam parity, decomposed e2e), per-stream timeline of one RHS
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
N=${NGPU:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/x13_pytest_multi_n$N.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/x13_bench_n$N.json 2> gpurun_out/x13_bench_n$N.err
EULERB200_HALO=p2p timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-parity > gpurun_out/x13_bench_n${N}_p2p.json 2> gpurun_out/x13_bench_n${N}_p2p.err
nproc > gpurun_out/x13_nproc_n$N.txt; free -g >> gpurun_out/x13_nproc_n$N.txt; nvidia-smi topo -m >> gpurun_out/x13_nproc_n$N.txt 2>&1
echo done > gpurun_out/x13_done_n$N.txt
