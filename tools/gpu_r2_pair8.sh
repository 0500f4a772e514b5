#!/bin/bash
# CTA pairs with ONE issuer thread and one multicast commit per row (ESRP_PAIR_SINGLE=1): parity, then A/B on the engine
mkdir -p gpurun_out
ESRP_PAIR_SINGLE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "cta_pairs" --tb=short 2>&1 | tail -4
for v in "ESRP_X=0" "ESRP_PAIR=1" "ESRP_PAIR=1 ESRP_PAIR_SINGLE=1" "ESRP_X=0" "ESRP_PAIR=1 ESRP_PAIR_SINGLE=1"; do
  env $v timeout 200 python tools/bench_fwd.py 30
done 2> gpurun_out/pair8_err.log | tee gpurun_out/pair8_ab.jsonl
tail -3 gpurun_out/pair8_err.log
