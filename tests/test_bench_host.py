"""Host-side checks of bench.py that need no GPU: it has no retry path any more (a device fault must fail the run),
and its source hash (which ties a committed ncu capture to a build) is stable."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_has_no_retry_supervisor():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "def supervise" not in src and "--worker" not in src
    assert not re.search(r"ESRP_ROW_ALT", src), "bench.py must not switch kernel protocols behind the measurement"


def test_csrc_hash_is_deterministic_and_tracks_sources(tmp_path):
    b = _load_bench()
    h1, h2 = b.csrc_hash(), b.csrc_hash()
    assert h1 == h2 and len(h1) == 64
