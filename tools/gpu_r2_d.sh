#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/d_diag.jsonl
for SHAPE in "3 37 100 1" "16 128 128 2"; do
  timeout 120 python tools/chain_diag.py $SHAPE >> gpurun_out/d_diag.jsonl 2>> gpurun_out/d_err.log; echo "rc=$?"
done
cat gpurun_out/d_diag.jsonl; tail -5 gpurun_out/d_err.log
: > gpurun_out/d_ab.jsonl
for v in ESRP_X=1 ESRP_CHAIN_DBG=2 ESRP_CHAIN_DBG=6 ESRP_CHAIN_DEP_ALL=1 ESRP_NO_CHAIN=1; do
  env $v timeout 200 python tools/bench_fwd.py 20 >> gpurun_out/d_ab.jsonl 2>> gpurun_out/d_err.log; echo "rc=$? $v"
done
cat gpurun_out/d_ab.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chain or config2" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_chain -c 1 -f -o gpurun_out/d_prof python tools/profile_step.py --nb 2 > gpurun_out/d_ncu_full.log 2>&1; echo "ncu full rc=$?"
