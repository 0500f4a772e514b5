#!/bin/bash
mkdir -p gpurun_out
for d in 2 10 258 266; do
ESRP_CHAIN_DBG=$d timeout 200 python tools/chain_trace.py 4 > gpurun_out/f_trace_dbg$d.json 2>> gpurun_out/f_err.log; echo "rc=$? dbg=$d"
python - <<PY
import json
d=json.load(open("gpurun_out/f_trace_dbg$d.json"))["phases"]
print("dbg=$d", {k:(v["issue(first->last MMA)"], v["phase(flag->flag)"], v["epi_tail(last MMA->flag)"]) for k,v in d.items()})
PY
done
: > gpurun_out/f_ab.jsonl
for v in ESRP_CHAIN_DBG=2 ESRP_CHAIN_DBG=10 ESRP_CHAIN_DBG=258 ESRP_CHAIN_DBG=266; do
  env $v timeout 200 python tools/bench_fwd.py 20 >> gpurun_out/f_ab.jsonl 2>> gpurun_out/f_err.log; echo "rc=$? $v"
done
cat gpurun_out/f_ab.jsonl; tail -3 gpurun_out/f_err.log
