#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_train.py --batch 32 --steps 8 > gpurun_out/ac_train.log 2>&1; tail -1 gpurun_out/ac_train.log
timeout 300 python tools/bench_train.py --batch 32 --steps 8 --perceptual > gpurun_out/ac_train_perceptual.log 2>&1; tail -2 gpurun_out/ac_train_perceptual.log
