#!/bin/bash
mkdir -p gpurun_out
for v in ESRP_NO_GRAPH=1 ESRP_X=0; do
  env $v timeout 200 python tools/bench_fwd_graph.py 30
done 2> gpurun_out/x_err.log | tee gpurun_out/x_graph_ab2.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench.py -x -q 2>&1 | tail -3
tail -3 gpurun_out/x_err.log
