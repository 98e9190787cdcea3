#!/bin/bash
# First GPU call of the next round (about 6 min of box time): everything that was written after the
# last GPU session of round 1 gets its measurement here.  Results land in gpurun_out/.
#   gpurun --timeout 600 -- 'bash tools/gpu_round2_first.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
# 1. the late additions: hook-assigned forcing (GW), boundary-heavy instantiation (AG), drop-in with a varying hook
timeout 200 python -m pytest tests/test_gpu_zz_forcing.py -q > gpurun_out/r2_zz_tests.log 2>&1
# 2. AG against the default on the configurations where boundary tiles are many
for k in 0 1; do
  EULERB200_KERNEL=$k timeout 120 python tools/configs_bench.py > gpurun_out/r2_configs_kernel$k.log 2>&1
done
# 3. 512^3 / NVAR 15 must not care (15 % boundary tiles: default instantiation either way)
for k in 0 1; do
  EULERB200_KERNEL=$k timeout 60 python tools/tune.py --n 512 512 512 --nchem 10 --variants 1 --steps 5 >> gpurun_out/r2_512_kernel.log 2>&1
done
# 3b. z-segment count: CTAs per launch aimed for (default 5920 -> 8 segments of 64 planes at 512^3)
for ctas in 2960 3996 4440 8880; do
  EULERB200_CTAS=$ctas timeout 60 python tools/tune.py --n 512 512 512 --nchem 10 --variants 1 --steps 5 2>&1 | sed "s/^/CTAS=$ctas /" >> gpurun_out/r2_512_ctas.log
done
# 4. race / memory check of the kernels on a small case (compute-sanitizer is in the CUDA toolkit)
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -k "test_feuler_matches_oracle and 16" > gpurun_out/r2_racecheck.log 2>&1
echo done > gpurun_out/r2_first_done.txt
