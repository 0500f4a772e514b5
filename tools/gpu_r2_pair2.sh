#!/bin/bash
# cta_group::1 vs ::2 MMA cost per N, then the whole GPU suite on the final sources (default path = split MMAs at the ring end,
# CTA pairs opt-in), then one more A/B of the engine with ESRP_PAIR=1
mkdir -p gpurun_out
timeout 120 ./tools/ubench_mma2 148 200 | tee gpurun_out/pair2_ubench_mma2.jsonl
timeout 60 ./tools/ubench_mma2 2 200 | tee gpurun_out/pair2_ubench_mma2_one_pair.jsonl
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/pair2_pytest_gpu.log
for i in 1 2; do
  timeout 200 python tools/bench_fwd.py 30
  ESRP_PAIR=1 timeout 200 python tools/bench_fwd.py 30
done 2> gpurun_out/pair2_err.log | tee gpurun_out/pair2_ab.jsonl
tail -3 gpurun_out/pair2_err.log
