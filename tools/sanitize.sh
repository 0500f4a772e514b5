#!/bin/bash
# compute-sanitizer memcheck over one small case of each path (forward, generator fwd+bwd, discriminator fwd+bwd).
mkdir -p gpurun_out
export ESRP_NO_PDL=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_train.py -q -x -k "32-1-shape1 or frozen" > gpurun_out/sanitize_g.log 2>&1; echo "G rc=$?"
tail -4 gpurun_out/sanitize_g.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_train_d.py -q -x -k "frozen or eval_mode" > gpurun_out/sanitize_d.log 2>&1; echo "D rc=$?"
tail -4 gpurun_out/sanitize_d.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python __graft_entry__.py smoke > gpurun_out/sanitize_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/sanitize_smoke.log
