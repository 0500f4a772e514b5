#!/bin/bash
# ncu --set full capture of the backward-only kernels (tcgen05 wgrad, conv1x1 backward, bias column sums) inside one
# RDB-sized training step (nb=1, 16x128x128), for profiles/.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"wgrad_tc_kernel|conv1x1_bwd|colsum" -s 10 -c 3 -f -o gpurun_out/prof_wgrad python tools/profile_step.py --nb 1 --bwd --train > gpurun_out/ncu_wgrad.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/prof_wgrad.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_raw.csv 2>/dev/null
cuobjdump -sass esrganplus_b200/libesrp.so 2>/dev/null | grep -E "UTCHMMA|UTCBAR|UTMALDG|UTCQMMA|UTCMMA" | sort | uniq -c | sort -rn | head -20 > gpurun_out/sass_tcgen05_mnemonics.txt
echo done
