#!/bin/bash
# 2-GPU check of the bench (flat-gradient all-reduce path) + strong-scaling train leg
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/t_bench2.json 2> gpurun_out/t_bench2.err
echo "rc=$?"; tail -c 3000 gpurun_out/t_bench2.json; tail -5 gpurun_out/t_bench2.err
