#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_poison.py 16 > gpurun_out/v_poison.log 2>&1; tail -60 gpurun_out/v_poison.log
