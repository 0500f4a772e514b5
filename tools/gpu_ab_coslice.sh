mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "coscheduled or rrdbnet or config2 or noise or tiled" > gpurun_out/pytest_slices.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_slices.log
ESRP_NO_COSLICE=1 timeout 300 python bench.py > gpurun_out/bench_nocoslice.log 2>&1; echo "bench(no coslice) rc=$?"
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_nocoslice.log","gpurun_out/bench.log"):
    try:
        l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["train"]["value"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
