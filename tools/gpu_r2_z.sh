#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python bench.py --no-extras --no-train > gpurun_out/z_bench_$i.json 2> gpurun_out/z_bench_err_$i.log
  python - <<PY
import json
d = json.loads(open("gpurun_out/z_bench_$i.json").read().strip().splitlines()[-1])
print($i, round(d["ms_per_step"], 3), d["e2e"], round(d["e2e_uint8"]["ms_per_step"], 3), d["roofline"]["frac"], d["roofline"]["traffic"], d["clocks"]["sm_mhz"])
PY
done
