#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/b_diag.jsonl
run() { env "$@" timeout 120 python tools/chain_diag.py $SHAPE >> gpurun_out/b_diag.jsonl 2>> gpurun_out/b_err.log; echo "rc=$? $@"; }
for SHAPE in "3 37 100 1" "2 128 128 1"; do
run ESRP_X=default
run ESRP_CHAIN_MAX=1
run ESRP_CHAIN_MAX=2
run ESRP_CHAIN_DEP_ALL=1
run ESRP_CHAIN_DBG=1
done
SHAPE="16 128 128 2"
run ESRP_X=default
run ESRP_CHAIN_DEP_ALL=1
cat gpurun_out/b_diag.jsonl; tail -5 gpurun_out/b_err.log
: > gpurun_out/b_ab.jsonl
for v in ESRP_X=1 ESRP_CHAIN_DBG=1 ESRP_CHAIN_DBG=2 ESRP_CHAIN_DBG=4 ESRP_CHAIN_DBG=7 ESRP_CHAIN_MAX=1 ESRP_CHAIN_MAX=5 ESRP_NO_CHAIN=1; do
  env $v timeout 200 python tools/bench_fwd.py 20 >> gpurun_out/b_ab.jsonl 2>> gpurun_out/b_err.log; echo "rc=$? $v"
done
cat gpurun_out/b_ab.jsonl
