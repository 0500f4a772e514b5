#!/bin/bash
# N-GPU check of the bench: bash tools/gpu_r2_t.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/t_bench$N.json 2> gpurun_out/t_bench$N.err
echo "rc=$?"; tail -c 2500 gpurun_out/t_bench$N.json; tail -3 gpurun_out/t_bench$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/t_bench_ref$N.json 2> gpurun_out/t_bench_ref$N.err
echo "ref rc=$?"; tail -c 400 gpurun_out/t_bench_ref$N.json
