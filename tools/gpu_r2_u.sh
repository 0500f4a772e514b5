#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 16" "16 16" "4 2" ; do
  echo "== $cfg"; timeout 200 python tools/diag_train_twice.py $cfg 2>&1 | grep -v "^[0-9]* [3-9] \|^[0-9]* 10 " | tail -14
done > gpurun_out/u_train_a.log 2>&1
for v in ESRP_D_GRAPH=0 ESRP_NO_GRAPH=1 DIAG_TORCH_SOLVER=1; do
  echo "== $v 32 16"; env $v timeout 200 python tools/diag_train_twice.py 32 16 2>&1 | grep -v "^[0-9]* [3-9] \|^[0-9]* 10 " | tail -7
done > gpurun_out/u_train_b.log 2>&1
cat gpurun_out/u_train_a.log gpurun_out/u_train_b.log
