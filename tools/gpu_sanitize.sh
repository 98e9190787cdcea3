#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (small grids): racecheck (shared-memory hazards of the
# row rendezvous / flux exchange), synccheck (named barriers), memcheck (ghost maps, halo slabs, sub-box launches)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
SEL='test_feuler_matches_oracle or test_row_synchronisation_modes_agree_bitwise or test_illegal_state'
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
     python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_forcing.py -x -q -k "$SEL or forcing or boundary" > gpurun_out/x15_sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/x15_sanitize_$tool.log
done
for pair in 0 1; do
  EULERB200_PAIR=$pair timeout 600 compute-sanitizer --tool racecheck --target-processes all --print-limit 20 \
     python -m pytest tests/test_gpu_parity.py -x -q -k "test_feuler_matches_oracle" > gpurun_out/x15_sanitize_racecheck_pair$pair.log 2>&1
  echo "rc=$?" >> gpurun_out/x15_sanitize_racecheck_pair$pair.log
done
EULERB200_SPLIT=1 timeout 600 compute-sanitizer --tool racecheck --target-processes all --print-limit 20 \
     python -m pytest tests/test_gpu_parity.py -x -q -k "test_feuler_matches_oracle" > gpurun_out/x15_sanitize_racecheck_split.log 2>&1
echo "rc=$?" >> gpurun_out/x15_sanitize_racecheck_split.log
echo done > gpurun_out/x15_done.txt
