#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_row -s 6 -c 6 -f -o gpurun_out/ncu_full_final python tools/profile_step.py > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/ncu_full_final.ncu-rep --page raw --csv > gpurun_out/ncu_full_final_raw.csv 2>/dev/null; echo "raw rc=$?"
rm -f gpurun_out/ncu_full_final.ncu-rep
