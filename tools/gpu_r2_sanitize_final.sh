#!/bin/bash
# compute-sanitizer over the round's new code paths on the final sources
mkdir -p gpurun_out
export ESRP_NO_PDL=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_vgg.py -x -q -k "elementwise or (other_shapes and 16) or rejects" > gpurun_out/sz_memcheck_vgg.log 2>&1; echo "memcheck vgg rc=$?"; tail -3 gpurun_out/sz_memcheck_vgg.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_data.py tests/test_gpu_solver.py -x -q -k "not gan_step" > gpurun_out/sz_memcheck_data_solver.log 2>&1; echo "memcheck data+solver rc=$?"; tail -3 gpurun_out/sz_memcheck_data_solver.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(coscheduled and shape0) or k_valid or config1 or repeated_forwards" > gpurun_out/sz_memcheck_row.log 2>&1; echo "memcheck row rc=$?"; tail -3 gpurun_out/sz_memcheck_row.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(coscheduled and shape0) or k_valid" > gpurun_out/sz_racecheck_row_chunk_bars.log 2>&1; echo "racecheck row (chunk barriers) rc=$?"; tail -3 gpurun_out/sz_racecheck_row_chunk_bars.log
