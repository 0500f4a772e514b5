#!/bin/bash
# where the CTA-pair launches lose: with / without programmatic dependent launch, and per K-chunk count
mkdir -p gpurun_out
for v in "ESRP_X=0" "ESRP_PAIR=1" "ESRP_NO_PDL=1" "ESRP_NO_PDL=1 ESRP_PAIR=1" "ESRP_PAIR=1 ESRP_PAIR_CHUNKS=2" "ESRP_PAIR=1 ESRP_PAIR_CHUNKS=4" "ESRP_PAIR=1 ESRP_PAIR_CHUNKS=8" "ESRP_X=0"; do
  env $v timeout 200 python tools/bench_fwd.py 30
done 2> gpurun_out/pair3_err.log | tee gpurun_out/pair3_ab.jsonl
tail -3 gpurun_out/pair3_err.log
