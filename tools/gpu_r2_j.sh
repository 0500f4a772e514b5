#!/bin/bash
mkdir -p gpurun_out
for d in 266 0; do
ESRP_CHAIN_DBG=$d timeout 200 python tools/chain_trace.py 4 > gpurun_out/j_trace_dbg$d.json 2>> gpurun_out/j_err.log; echo "rc=$? dbg=$d"
python - <<PY
import json
d=json.load(open("gpurun_out/j_trace_dbg$d.json"))["phases"]
for k,v in d.items(): print("dbg=$d", k, {a[6:]:b for a,b in v.items() if a.startswith("row6")})
PY
done
tail -3 gpurun_out/j_err.log
