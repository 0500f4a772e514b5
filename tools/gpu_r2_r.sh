#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solver.py -q -x -k flat_adam 2>&1 | tail -30 > gpurun_out/r_adam.log; cat gpurun_out/r_adam.log
timeout 300 python tools/prof_host.py > gpurun_out/r_prof_host.log 2>&1; head -75 gpurun_out/r_prof_host.log
export ESRP_NO_PDL=1
timeout 600 compute-sanitizer --tool synccheck --print-limit 200 --show-backtrace device python -m pytest tests/test_gpu_parity.py -x -q -k "coscheduled and shape0 and False" > gpurun_out/r_sync_row.log 2>&1
grep "Device Frame: void\|located at" gpurun_out/r_sync_row.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -6; tail -2 gpurun_out/r_sync_row.log
