#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/t_diag.jsonl
for m in torch native; do timeout 120 python tools/diag_solver.py $m >> gpurun_out/t_diag.jsonl 2>> gpurun_out/t_err.log; done
cat gpurun_out/t_diag.jsonl; tail -3 gpurun_out/t_err.log
