#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/e_ab.jsonl
for v in ESRP_X=1 ESRP_CHAIN_DBG=48 ESRP_CHAIN_DBG=64 ESRP_CHAIN_DBG=96 ESRP_CHAIN_DBG=128 ESRP_CHAIN_DBG=70 ESRP_NO_CHAIN=1; do
  env $v timeout 200 python tools/bench_fwd.py 20 >> gpurun_out/e_ab.jsonl 2>> gpurun_out/e_err.log; echo "rc=$? $v"
done
cat gpurun_out/e_ab.jsonl; tail -3 gpurun_out/e_err.log
