# quick A/B of timing switches on config 2 (forward only)
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>gpurun_out/ab_err.log
ESRP_TMAP_PROMO256=1 timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
ESRP_NO_PDL=1 timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
timeout 120 python tools/bench_fwd.py >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
cat gpurun_out/ab.jsonl
