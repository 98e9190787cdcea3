#!/bin/bash
# full ncu capture of the fused RHS kernel at 512^3 / NVAR 15 (one launch); raw + source pages exported on the box
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
tag=${NCU_TAG:-x11_fused}
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rhs_fused -s 3 -c 1 -f -o /tmp/$tag \
   python tools/tune2.py --n 512 512 512 --nchem ${NCU_NCHEM:-10} --steps 1 --env "${NCU_ENV:-}" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page source --csv > gpurun_out/$tag.source.csv 2>/dev/null
python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/$tag.txt 2>&1
cp /tmp/$tag.ncu-rep gpurun_out/ 2>/dev/null
echo done > gpurun_out/${tag}_done.txt
