#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vgg.py -x -q -s 2>&1 | tail -40 > gpurun_out/aa_vgg.log; cat gpurun_out/aa_vgg.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_d.py tests/test_gpu_backward.py -x -q 2>&1 | tail -5
