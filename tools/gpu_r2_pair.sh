#!/bin/bash
# CTA-pair row kernel (ESRP_VARIANT_PAIR): parity first, then same-box A/B against ESRP_PAIR=0, then the parity file
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cta_pairs" --tb=short 2>&1 | tail -40 | tee gpurun_out/pair_pytest_pairs.log
for i in 1 2 3; do
  ESRP_PAIR=0 timeout 200 python tools/bench_fwd.py 30
  ESRP_PAIR=1 timeout 200 python tools/bench_fwd.py 30
done 2> gpurun_out/pair_err.log | tee gpurun_out/pair_ab.jsonl
tail -3 gpurun_out/pair_err.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=line 2>&1 | tail -15 | tee gpurun_out/pair_pytest_parity.log
