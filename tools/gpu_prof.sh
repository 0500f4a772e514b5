# ncu launch list of one config-2 forward + full capture of one dense block's conv launches
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_c2.csv python tools/profile_step.py > gpurun_out/ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_row -s 6 -c 6 -f -o gpurun_out/prof python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null; echo "raw rc=$?"
