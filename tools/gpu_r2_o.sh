#!/bin/bash
mkdir -p gpurun_out
export ESRP_NO_PDL=1
run() { name=$1; shift; timeout 600 compute-sanitizer --tool synccheck --print-limit 200000 --show-backtrace no "$@" > gpurun_out/o_sync_$name.log 2>&1; echo "$name rc=$?";
  grep -A3 "Barrier error\|error detected" gpurun_out/o_sync_$name.log | grep -v "^--" | sed 's/thread ([0-9]*,0,0) in block ([0-9]*,0,0)/thread T in block B/' | sort | uniq -c | sort -rn | head -12; tail -3 gpurun_out/o_sync_$name.log; }
run row_alt2 env ESRP_ROW_ALT=2 python -m pytest tests/test_gpu_parity.py -x -q -k "coscheduled and shape0 and False"
run row_alt0 env ESRP_ROW_ALT=0 python -m pytest tests/test_gpu_parity.py -x -q -k "coscheduled and shape0 and False"
run chain python -m pytest tests/test_gpu_parity.py -x -q -k "chain_matches and shape4"
run tile python -m pytest tests/test_gpu_parity.py -x -q -k "rrdbnet_config1"
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_data.py tests/test_gpu_train.py tests/test_gpu_train_d.py -q -x 2>&1 | tail -25 > gpurun_out/o_pytest_new.log; cat gpurun_out/o_pytest_new.log
