#!/bin/bash
# round-2 late A/B: leaner arithmetic (third-order reciprocal in weno5, 0.5 of the split folded into 1/dx,
# shared state projection in fluid_face, negated species divergence) against the library of commit caa39a0
# (copied to libeulerb200_prev.so) on one box, then the GPU test suite on the new default
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
T=${TAG:-x40}
P=$PWD/sundials-manyvector-demo_b200
for lib in libeulerb200_prev.so libeulerb200.so libeulerb200_projclassic.so; do
  [ -f $P/$lib ] || continue
  echo "== $lib" >> gpurun_out/${T}_tune.log
  EULERB200_LIB=$P/$lib timeout 120 python tools/tune2.py --n 512 512 512 --nchem 10 --steps 5 --env "" >> gpurun_out/${T}_tune.log 2>&1
done
for lib in libeulerb200_prev.so libeulerb200.so; do
  echo "== $lib nchem=0" >> gpurun_out/${T}_tune.log
  EULERB200_LIB=$P/$lib timeout 120 python tools/tune2.py --n 512 512 512 --nchem 0 --steps 5 --env "" >> gpurun_out/${T}_tune.log 2>&1
done
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1
echo done > gpurun_out/${T}_done.txt
