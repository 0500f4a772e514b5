#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f3_smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/f3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f3_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/f3_launches_c2.csv python tools/profile_step.py > gpurun_out/f3_ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
