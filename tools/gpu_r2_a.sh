#!/bin/bash
# Round 2, GPU call A: first hardware exposure of the persistent conv chain (csrc/conv3x3_chain.cuh).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chain" > gpurun_out/a_pytest_chain.log 2>&1; echo "pytest chain rc=$?"
tail -15 gpurun_out/a_pytest_chain.log
: > gpurun_out/a_ab.jsonl
for rep in 1 2; do
  timeout 200 python tools/bench_fwd.py 30 >> gpurun_out/a_ab.jsonl 2>> gpurun_out/a_ab_err.log; echo "chain rc=$?"
  ESRP_NO_CHAIN=1 timeout 200 python tools/bench_fwd.py 30 >> gpurun_out/a_ab.jsonl 2>> gpurun_out/a_ab_err.log; echo "nochain rc=$?"
done
cat gpurun_out/a_ab.jsonl; tail -5 gpurun_out/a_ab_err.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "config2 or config1 or nb23 or uint8 or noise" > gpurun_out/a_pytest_net.log 2>&1; echo "pytest net rc=$?"
tail -5 gpurun_out/a_pytest_net.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/a_launches_c2.csv python tools/profile_step.py > gpurun_out/a_ncu_list.log 2>&1; echo "ncu list rc=$?"
tail -3 gpurun_out/a_ncu_list.log
