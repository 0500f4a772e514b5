#!/bin/bash
# re-validation of the final build after the single-issuer pair switch was added: whole GPU suite, ncu launch list of a
# config-2 forward (summarised on the box so that bench.py finds a capture with the hash of THIS build), short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/f7_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/f7_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/f7_launches_c2.csv python tools/profile_step.py > gpurun_out/f7_ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
python tools/summarize_ncu.py list gpurun_out/f7_launches_c2.csv "config 2 forward (16x128x128 LR, nb=23), final sources of round 2" > gpurun_out/f7_ncu_launch_summary_config2.json \
  && cp gpurun_out/f7_ncu_launch_summary_config2.json profiles/r02_ncu_launch_summary_config2.json; echo "summary rc=$?"
timeout 300 python bench.py --no-train --no-extras > gpurun_out/f7_bench.json 2> gpurun_out/f7_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/f7_bench.json
