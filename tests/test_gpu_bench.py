"""bench.py end to end on the GPU box: one short run must exit 0 and print the contract's JSON line — there is no retry
supervisor any more, so a device fault in any inference leg fails this test."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_runs_clean_without_retry(cuda_dev):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-train"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["attempts"] == 1 and "retry_env" not in d
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "tensor" and rf["achieved"] > 0 and 0 < rf["frac"] < 1.2 and rf["duration_ms"] > 0
    assert rf["duration_ms"] < d["ms_per_step"], "the trunk's dense-block convs are a part of the step"
    assert d["chain"] and "error" not in d["chain"], d["chain"]
    assert d["tiled"] and "error" not in d["tiled"], d["tiled"]
    assert d["gpu_library_baseline"] and "error" not in d["gpu_library_baseline"], d["gpu_library_baseline"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
