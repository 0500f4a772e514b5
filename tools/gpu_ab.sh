#!/bin/bash
# Quick A/B of the row kernel's timing switches on config 2 (forward only, one process per setting: the switches are read
# once per process).  Each line of gpurun_out/ab.jsonl names its environment.  See DESIGN.md section 5 for the switches.
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
run() { env "$@" timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>> gpurun_out/ab_err.log; }
: > gpurun_out/ab_err.log
run ESRP_AB=default
run ESRP_ROW_ALT=0
run ESRP_NO_HALF_CHUNK=1
run ESRP_NO_QUAD=1
run ESRP_NO_PLANAR=1
run ESRP_NO_COSLICE=1
run ESRP_TMAP_PROMO256=1
run ESRP_NO_PDL=1
cat gpurun_out/ab.jsonl
