#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_batch.py > gpurun_out/m_batch.jsonl 2> gpurun_out/m_err.log; echo "rc=$?"; cat gpurun_out/m_batch.jsonl; tail -3 gpurun_out/m_err.log
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity.py::test_config2_full_size_properties > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/m_pytest.log
