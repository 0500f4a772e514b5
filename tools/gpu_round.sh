#!/bin/bash
# Standard GPU visit: smoke, GPU tests, bench (both arms), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -2 gpurun_out/bench.log
if [ "$1" != "noprof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_tc -c 7 -f -o gpurun_out/prof python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
tail -1 gpurun_out/bench_ref.log
