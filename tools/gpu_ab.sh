# all GPU tests + quick A/B of timing switches on config 2 (forward only)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/ab.jsonl
timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>gpurun_out/ab_err.log
ESRP_NO_HALF_CHUNK=1 timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
ESRP_ROW_ALT=0 timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
cat gpurun_out/ab.jsonl
