#!/bin/bash
# GPU visit: all GPU tests, bench (both arms), ncu launch lists (config 2 forward; RDB fwd+bwd micro = config 5), full capture.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_c2.csv python tools/profile_step.py > gpurun_out/ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_rdb_train.csv python tools/profile_step.py --nb 1 --bwd --train > gpurun_out/ncu_list_rdb.log 2>&1; echo "ncu list rdb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_row -s 6 -c 6 -f -o gpurun_out/prof python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null; echo "raw rc=$?"
