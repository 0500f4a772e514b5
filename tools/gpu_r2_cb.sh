#!/bin/bash
mkdir -p gpurun_out
for v in "ESRP_ROW_ALT=2 ESRP_CHUNK_BARS=0" "ESRP_ROW_ALT=2 ESRP_CHUNK_BARS=1" "ESRP_ROW_ALT=3 ESRP_CHUNK_BARS=1" "ESRP_ROW_ALT=3 ESRP_CHUNK_BARS=2" "ESRP_ROW_ALT=2 ESRP_CHUNK_BARS=0" "ESRP_ROW_ALT=2 ESRP_CHUNK_BARS=1" "ESRP_ROW_ALT=3 ESRP_CHUNK_BARS=1" "ESRP_ROW_ALT=3 ESRP_CHUNK_BARS=2"; do
  env $v timeout 200 python tools/bench_fwd.py 30
done 2> gpurun_out/cb_err.log | tee gpurun_out/cb_ab2.jsonl
tail -2 gpurun_out/cb_err.log
ESRP_ROW_ALT=3 ESRP_CHUNK_BARS=2 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
