"""bench.py at N = 1 runs the measurement in a worker process and repeats a failed attempt once (a device fault poisons
the CUDA context of the process it happens in).  Host logic only: the worker is faked."""
import importlib.util
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("fail_first", [False, True])
def test_supervisor_relays_the_worker_line_and_retries_once(monkeypatch, capsys, fail_first):
    import torch
    from esrganplus_b200 import _lib
    bench = _load_bench()
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(_lib, "load", lambda: None)
    calls = []

    def fake_run(cmd, env=None, stdout=None, text=None):
        calls.append((cmd, env.get("ESRP_ROW_ALT")))
        assert cmd[-1] == "--worker" and cmd[1].endswith("bench.py")
        if fail_first and len(calls) == 1:
            return types.SimpleNamespace(returncode=1, stdout="Traceback ...\n")
        return types.SimpleNamespace(returncode=0, stdout='note\n{"metric": "m", "value": 1.5}\n')

    monkeypatch.setattr(subprocess, "run", fake_run)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "3"])
    monkeypatch.delenv("ESRP_ROW_ALT", raising=False)
    bench.supervise()
    out = capsys.readouterr()
    line = json.loads([l for l in out.out.splitlines() if l.startswith("{")][-1])
    assert line["metric"] == "m" and line["value"] == 1.5
    assert line["attempts"] == (2 if fail_first else 1)
    assert calls[0][0][2:4] == ["--steps", "3"] and calls[0][1] is None
    if fail_first:
        assert calls[1][1] == "0" and line["retry_env"] == {"ESRP_ROW_ALT": "0"} and "attempt 1 failed" in out.err
    else:
        assert len(calls) == 1 and "retry_env" not in line


def test_supervisor_gives_up_after_two_failures(monkeypatch):
    import torch
    from esrganplus_b200 import _lib
    bench = _load_bench()
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(_lib, "load", lambda: None)
    monkeypatch.setattr(subprocess, "run", lambda *a, **k: types.SimpleNamespace(returncode=1, stdout=""))
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    with pytest.raises(SystemExit):
        bench.supervise()
