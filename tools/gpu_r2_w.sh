#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train_d.py -x -q -k "first_pass" 2>&1 | tail -5
timeout 200 python tools/diag_train_twice.py 32 16 2>&1 | grep -v "^[0-9]* [3-9] \|^[0-9]* 10 " | tail -8
timeout 200 python tools/diag_train_twice.py 4 2 2>&1 | grep -v "^[0-9]* [3-9] \|^[0-9]* 10 " | tail -4
