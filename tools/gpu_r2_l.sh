#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/l_ab.jsonl
for v in ESRP_X=1 ESRP_CHAIN_DBG=2 ESRP_NO_CHAIN=1 ESRP_X=2 ESRP_NO_CHAIN=1; do
  env $v timeout 200 python tools/bench_fwd.py 30 >> gpurun_out/l_ab.jsonl 2>> gpurun_out/l_err.log; echo "rc=$? $v"
done
cat gpurun_out/l_ab.jsonl; tail -3 gpurun_out/l_err.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chain or config2" > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/l_pytest.log
