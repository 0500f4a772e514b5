#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/s_pytest_full.log; cat gpurun_out/s_pytest_full.log
: > gpurun_out/s_train.log
for b in 32 4; do timeout 600 python tools/bench_train.py --batch $b 2>&1 | tail -1 >> gpurun_out/s_train.log; done; cat gpurun_out/s_train.log
timeout 900 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench_err.log; echo "bench rc=$?"; cat gpurun_out/s_bench.json | cut -c1-6000
