#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/c_diag.jsonl
run() { env "$@" timeout 120 python tools/chain_diag.py $SHAPE >> gpurun_out/d_diag.jsonl 2>> gpurun_out/c_err.log; echo "rc=$? $@"; }
for SHAPE in "3 37 100 1" "2 128 128 1" "16 128 128 2"; do
run ESRP_X=default
done
SHAPE="2 128 128 1"
run ESRP_ROW_ALT=0
cat gpurun_out/c_diag.jsonl; tail -5 gpurun_out/c_err.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_chain -c 1 -f -o gpurun_out/c_prof python tools/profile_step.py --nb 2 > gpurun_out/c_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/c_ncu_full.log; ls -la gpurun_out/c_prof.ncu-rep
