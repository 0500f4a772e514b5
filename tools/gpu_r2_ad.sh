#!/bin/bash
# Full validation of the current state: smoke, all GPU tests, ncu launch lists + full capture, bench (both arms)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/ad_smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ad_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/ad_launches_c2.csv python tools/profile_step.py > gpurun_out/ad_ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/ad_launches_train_crop32.csv python tools/profile_step.py --batch 32 --tile 32 --bwd --train > gpurun_out/ad_ncu_list_train.log 2>&1; echo "ncu list train rc=$?"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/ad_launches_d.csv python tools/profile_d.py > gpurun_out/ad_ncu_list_d.log 2>&1; echo "ncu list d rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_row -s 6 -c 6 -f -o gpurun_out/ad_prof python tools/profile_step.py > gpurun_out/ad_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/ad_prof.ncu-rep --page raw --csv > gpurun_out/ad_prof_raw.csv 2>/dev/null; echo "raw rc=$?"
rm -f gpurun_out/ad_prof.ncu-rep
timeout 900 python bench.py > gpurun_out/ad_bench.json 2> gpurun_out/ad_bench_err.log; echo "bench rc=$?"; tail -c 1500 gpurun_out/ad_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ad_bench_ref.json 2> gpurun_out/ad_bench_ref_err.log; echo "ref rc=$?"; tail -c 600 gpurun_out/ad_bench_ref.json
