timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"wgrad_kernel|conv1x1_bwd" -s 4 -c 2 -f -o gpurun_out/prof_wgrad python tools/profile_step.py --nb 1 --bwd --train > gpurun_out/ncu_wgrad.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/prof_wgrad.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_wgrad.ncu-rep --page details --csv > gpurun_out/prof_wgrad_details.csv 2>/dev/null
echo done
