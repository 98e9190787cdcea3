#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
find . -name "*.so" -exec touch {} + ; touch sundials-manyvector-demo_b200/euler3d_b200 2>/dev/null
find oracle/_ref -type f -exec touch {} + 2>/dev/null
# full captures of the fluid-only (384) and species-only (640) launches and of the fused kernel at 512^3 / NVAR 15;
# raw + source pages exported here (the reports themselves are too big to bring back together)
timeout 600 ncu --set full --clock-control none -k regex:rhs_fused -s 6 -c 2 -f -o /tmp/x3_split_t640 \
   python tools/tune2.py --n 512 512 512 --nchem 10 --steps 1 --env "SPLIT=1 VARIANT_T=3" > gpurun_out/x3_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:rhs_fused -s 3 -c 1 -f -o /tmp/x3_fused \
   python tools/tune2.py --n 512 512 512 --nchem 10 --steps 1 --env "" > gpurun_out/x3_ncu_b.log 2>&1
for r in x3_split_t640 x3_fused; do
  ncu -i /tmp/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i /tmp/$r.ncu-rep --page source --csv > gpurun_out/$r.source.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/$r.ncu-rep > gpurun_out/$r.txt 2>&1
done
ls -la /tmp/*.ncu-rep > gpurun_out/x3_sizes.txt
echo done > gpurun_out/x3_done.txt
