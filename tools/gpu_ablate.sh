#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
out=gpurun_out/${ABLATE_OUT:-x8_ablate.log}
for m in ${ABLATE_MASKS:-0 31}; do
  echo "chemT copy:" >> $out; timeout 120 tools/ablate_$m 512 10 1 >> $out 2>&1
  echo "species-fastest vector:" >> $out; timeout 120 tools/ablate_$m 512 10 0 >> $out 2>&1
done
echo done > gpurun_out/x8_done.txt
