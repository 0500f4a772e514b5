#!/bin/bash
# robustness batch on the FINAL sources of round 2 (default path recompiled with the pair / shadow template switches off):
# NVML poller beside 16 forward processes, 3 bench processes and 2 full suites
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 20 > gpurun_out/s6_nvml_poll.log 2>&1 &
POLL=$!
: > gpurun_out/s6_stress.jsonl; : > gpurun_out/s6_stress_bench.jsonl; : > gpurun_out/s6_err.log
fa=0; for i in $(seq 1 16); do timeout 120 python tools/bench_fwd.py 30 >> gpurun_out/s6_stress.jsonl 2>> gpurun_out/s6_err.log || { fa=$((fa+1)); echo "fwd run $i failed"; }; done
echo "forward processes failed: $fa of 16"
fb=0; for i in 1 2 3; do timeout 300 python bench.py --no-train --no-extras --steps 5 --warmup 3 >> gpurun_out/s6_stress_bench.jsonl 2>> gpurun_out/s6_err.log || { fb=$((fb+1)); echo "bench run $i failed"; }; done
echo "bench processes failed: $fb of 3"
fc=0; for i in 1 2; do timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/s6_pytest_$i.log 2>&1 || { fc=$((fc+1)); echo "suite $i failed"; tail -5 gpurun_out/s6_pytest_$i.log; }; tail -1 gpurun_out/s6_pytest_$i.log; done
echo "full suites failed: $fc of 2"
kill $POLL
wc -l gpurun_out/s6_nvml_poll.log
sort -t, -k2 -n -r gpurun_out/s6_nvml_poll.log | head -2
