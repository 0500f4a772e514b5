#!/bin/bash
# same-box A/B of two library builds: esrganplus_b200/libesrp_base.so (before) vs libesrp.so (after)
mkdir -p gpurun_out
for i in 1 2 3; do
  ESRP_LIBRARY=$PWD/esrganplus_b200/libesrp_base.so timeout 200 python tools/bench_fwd.py 30
  timeout 200 python tools/bench_fwd.py 30
done 2> gpurun_out/ab_lib_err.log | tee gpurun_out/ab_lib.jsonl
tail -2 gpurun_out/ab_lib_err.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
