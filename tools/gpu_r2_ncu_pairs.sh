#!/bin/bash
# ncu --set full of the conv3-shaped launch (K = 128) on single CTAs (variant 8192) and on CTA pairs (24576)
mkdir -p gpurun_out
for v in 8192 24576; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_row_kernel -s 6 -c 2 -f -o gpurun_out/ncu_pair_v$v \
     python tools/trace_gap.py conv3 v=$v > gpurun_out/ncu_pair_v$v.log 2>&1; echo "ncu v=$v rc=$?"
  ncu -i gpurun_out/ncu_pair_v$v.ncu-rep --page raw --csv > gpurun_out/ncu_pair_v${v}_raw.csv 2>/dev/null; echo "raw rc=$?"
done
ls -la gpurun_out/ncu_pair_*
