#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/q_pytest_full.log; cat gpurun_out/q_pytest_full.log
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_d.py -q -rP -k "nb23 or reference_fixture or fp32_oracle" 2>&1 | grep -E "nb23|reference|fixture|fp32 oracle|rel_l2|passed|failed" | cut -c1-300 > gpurun_out/q_grad_metrics.log; cat gpurun_out/q_grad_metrics.log
export ESRP_NO_PDL=1
timeout 600 compute-sanitizer --tool synccheck --print-limit 2000 --show-backtrace device python -m pytest tests/test_gpu_parity.py -x -q -k "chain_matches and shape4" > gpurun_out/q_sync_chain.log 2>&1
grep "Device Frame.*kernel\|Device Frame: void" gpurun_out/q_sync_chain.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head; tail -2 gpurun_out/q_sync_chain.log
unset ESRP_NO_PDL
for g in 1 0; do ESRP_D_GRAPH=$g timeout 600 python tools/bench_train.py 2>&1 | tail -3 >> gpurun_out/q_train.log; done; for b in 4 32; do timeout 600 python tools/bench_train.py --batch $b 2>&1 | tail -2 >> gpurun_out/q_train.log; done; cat gpurun_out/q_train.log
