#!/bin/bash
# fast-build A/B of a kernel source change against the committed default library (same box)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
T=${TAG:-x30}
timeout 200 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "" > gpurun_out/${T}_tune_base.log 2>&1
export EULERB200_LIB=$PWD/sundials-manyvector-demo_b200/libeulerb200_fast.so
timeout 200 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" "" ${EXTRA_ENVS} > gpurun_out/${T}_tune_new.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "test_feuler_matches_oracle" > gpurun_out/${T}_pytest.log 2>&1
echo done > gpurun_out/${T}_done.txt
