#!/bin/bash
# last robustness batch of round 2 on the final sources: NVML poller beside 60 forward processes, 10 bench processes, 6 full suites
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 20 > gpurun_out/sl_nvml_poll.log 2>&1 &
POLL=$!
: > gpurun_out/sl_stress.jsonl; : > gpurun_out/sl_stress_bench.jsonl; : > gpurun_out/sl_err.log
fa=0; for i in $(seq 1 60); do timeout 120 python tools/bench_fwd.py 30 >> gpurun_out/sl_stress.jsonl 2>> gpurun_out/sl_err.log || { fa=$((fa+1)); echo "fwd run $i failed"; }; done
echo "forward processes failed: $fa of 60"
fb=0; for i in $(seq 1 10); do timeout 300 python bench.py --no-train --no-extras --steps 5 --warmup 3 >> gpurun_out/sl_stress_bench.jsonl 2>> gpurun_out/sl_err.log || { fb=$((fb+1)); echo "bench run $i failed"; }; done
echo "bench processes failed: $fb of 10"
fc=0; for i in 1 2 3 4 5 6; do timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/sl_pytest_$i.log 2>&1 || { fc=$((fc+1)); echo "suite $i failed"; tail -5 gpurun_out/sl_pytest_$i.log; }; tail -1 gpurun_out/sl_pytest_$i.log; done
echo "full suites failed: $fc of 6"
kill $POLL
wc -l gpurun_out/sl_nvml_poll.log
