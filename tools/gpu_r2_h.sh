#!/bin/bash
mkdir -p gpurun_out
for d in 2 266; do
ESRP_CHAIN_DBG=$d timeout 200 python tools/chain_trace.py 4 > gpurun_out/h_trace_dbg$d.json 2>> gpurun_out/h_err.log; echo "rc=$? dbg=$d"
python - <<PY
import json
d=json.load(open("gpurun_out/h_trace_dbg$d.json"))["phases"]
for k,v in d.items(): print("dbg=$d", k, {a:b for a,b in v.items() if a.startswith("row4") or a.startswith("issue")})
PY
done
tail -3 gpurun_out/h_err.log
