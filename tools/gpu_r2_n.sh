#!/bin/bash
# Round 2 robustness call: sanitizer synccheck / racecheck over the row kernel in both issuer
# protocols and over the conv chain; process-launch stress with an NVML poller beside it; repeated full GPU suites.
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 20 > gpurun_out/n_nvml_poll.log 2>&1 &
POLL=$!
echo "== full suite #1"; timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/n_pytest_1.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/n_pytest_1.log
echo "== bench"; timeout 900 python bench.py --no-train > gpurun_out/n_bench.log 2> gpurun_out/n_bench_err.log; echo "rc=$?"; tail -c 3000 gpurun_out/n_bench.log
: > gpurun_out/n_stress.jsonl
fail=0
echo "== stress A: 100 processes x 30 config-2 forwards"
for i in $(seq 1 100); do
  timeout 120 python tools/bench_fwd.py 30 >> gpurun_out/n_stress.jsonl 2>> gpurun_out/n_stress_err.log || { fail=$((fail+1)); echo "fwd run $i failed"; }
done
echo "stress A failures: $fail"
echo "== stress B: 30 processes of bench.py --no-train --no-extras --steps 5"
failb=0
for i in $(seq 1 30); do
  timeout 300 python bench.py --no-train --no-extras --steps 5 --warmup 3 >> gpurun_out/n_stress_bench.jsonl 2>> gpurun_out/n_stress_err.log || { failb=$((failb+1)); echo "bench run $i failed"; }
done
echo "stress B failures: $failb"
echo "== stress C: 6 more full suites"
failc=0
for i in 2 3 4 5 6 7; do
  timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/n_pytest_$i.log 2>&1 || { failc=$((failc+1)); echo "suite $i failed"; tail -5 gpurun_out/n_pytest_$i.log; }
done
echo "stress C failures: $failc"
kill $POLL
wc -l gpurun_out/n_nvml_poll.log
echo "== sanitizers"
export ESRP_NO_PDL=1
for tool in synccheck racecheck; do
  for alt in 2 0; do
    ESRP_ROW_ALT=$alt timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -x -q -k "rrdbnet_config1 or (coscheduled and shape0) or (chain_matches and shape4) or (chain_matches and shape0)" \
      > gpurun_out/n_sanitizer_${tool}_alt$alt.log 2>&1
    echo "$tool alt=$alt rc=$?"; tail -4 gpurun_out/n_sanitizer_${tool}_alt$alt.log
  done
done
