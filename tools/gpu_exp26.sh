#!/bin/bash
# N-GPU call: exchange-first (EULERB200_OVERLAP=0) vs interior/shell overlap (=1), NCCL and peer-store transports;
# multi-GPU parity tests
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
N=${NGPU:-2}
port=29520
for ov in ${OVS:-1 0}; do for halo in ${HALOS:-nccl p2p}; do
  port=$((port+1))
  EULERB200_OVERLAP=$ov EULERB200_HALO=$halo timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/x${TAG:-26}_bench_n${N}_ov${ov}_${halo}.json 2> gpurun_out/x${TAG:-26}_bench_n${N}_ov${ov}_${halo}.err
done; done
if [ "$N" = "2" ] && [ -z "$SKIP_PYTEST" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/x${TAG:-26}_pytest_multi_n$N.log 2>&1
fi
echo done > gpurun_out/x${TAG:-26}_done_n$N.txt
