#!/bin/bash
# Final state of round 2 (CTA pairs: not for co-scheduled slices): the whole GPU suite, compute-sanitizer racecheck / memcheck over
# the pair kernels, the ncu launch list of a config-2 forward (summarised on the box so that bench.py finds a capture with
# the hash of THIS build), bench.py
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f5_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f5_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f5_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f5_pytest.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -k "cta_pairs and not shape4 and not engine" --tb=line \
   > gpurun_out/f5_sanitizer_racecheck_pairs.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/f5_sanitizer_racecheck_pairs.log | tail -3
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "cta_pairs and not shape4" --tb=line \
   > gpurun_out/f5_sanitizer_memcheck_pairs.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/f5_sanitizer_memcheck_pairs.log | tail -3
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/f5_launches_c2.csv python tools/profile_step.py > gpurun_out/f5_ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
python tools/summarize_ncu.py list gpurun_out/f5_launches_c2.csv "config 2 forward (16x128x128 LR, nb=23), final sources of round 2" > gpurun_out/f5_ncu_launch_summary_config2.json \
  && cp gpurun_out/f5_ncu_launch_summary_config2.json profiles/r02_ncu_launch_summary_config2.json; echo "summary rc=$?"
timeout 900 python bench.py > gpurun_out/f5_bench.json 2> gpurun_out/f5_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/f5_bench.json
