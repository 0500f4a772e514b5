#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/k_diag.jsonl
for SHAPE in "3 37 100 1" "16 128 128 2"; do
  timeout 120 python tools/chain_diag.py $SHAPE >> gpurun_out/k_diag.jsonl 2>> gpurun_out/k_err.log; echo "rc=$?"
done
cat gpurun_out/k_diag.jsonl; tail -5 gpurun_out/k_err.log
for d in 0 2 266; do
ESRP_CHAIN_DBG=$d timeout 200 python tools/chain_trace.py 4 > gpurun_out/k_trace_dbg$d.json 2>> gpurun_out/k_err.log; echo "rc=$? dbg=$d"
python - <<PY
import json
d=json.load(open("gpurun_out/k_trace_dbg$d.json"))["phases"]
print("dbg=$d", {k:(v["issue(first->last MMA)"], v["phase(flag->flag)"], v["epi_tail(last MMA->flag)"], v["flak_wait(2->3)"]) for k,v in d.items()})
PY
done
: > gpurun_out/k_ab.jsonl
for v in ESRP_X=1 ESRP_CHAIN_DBG=2 ESRP_CHAIN_DBG=266 ESRP_NO_CHAIN=1; do
  env $v timeout 200 python tools/bench_fwd.py 20 >> gpurun_out/k_ab.jsonl 2>> gpurun_out/k_err.log; echo "rc=$? $v"
done
cat gpurun_out/k_ab.jsonl; tail -3 gpurun_out/k_err.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chain or config2" > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/k_pytest.log
