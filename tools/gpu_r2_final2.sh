#!/bin/bash
# validation of the per-chunk landed barriers: full suite, stress with NVML poller, ncu list, bench
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f2_smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f2_pytest.log
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 20 > gpurun_out/f2_nvml_poll.log 2>&1 &
POLL=$!
: > gpurun_out/f2_stress.jsonl; : > gpurun_out/f2_err.log
fa=0; for i in $(seq 1 40); do timeout 120 python tools/bench_fwd.py 30 >> gpurun_out/f2_stress.jsonl 2>> gpurun_out/f2_err.log || { fa=$((fa+1)); echo "fwd run $i failed"; }; done
echo "forward processes failed: $fa of 40"
fb=0; for i in $(seq 1 8); do timeout 300 python bench.py --no-train --no-extras --steps 5 --warmup 3 >> gpurun_out/f2_stress_bench.jsonl 2>> gpurun_out/f2_err.log || { fb=$((fb+1)); echo "bench run $i failed"; }; done
echo "bench processes failed: $fb of 8"
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/f2_pytest_2.log 2>&1; echo "suite 2 rc=$?"; tail -1 gpurun_out/f2_pytest_2.log
kill $POLL
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/f2_launches_c2.csv python tools/profile_step.py > gpurun_out/f2_ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
for v in ESRP_CHUNK_BARS=0 ESRP_X=1 ESRP_CHUNK_BARS=0 ESRP_X=1; do env $v timeout 200 python tools/bench_fwd.py 30; done 2>> gpurun_out/f2_err.log | tee gpurun_out/f2_ab.jsonl
timeout 900 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench_err.log; echo "bench rc=$?"; tail -c 300 gpurun_out/f2_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f2_bench_ref.json 2> gpurun_out/f2_bench_ref_err.log; echo "ref rc=$?"
