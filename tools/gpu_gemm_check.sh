#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_gemm_tf32.py -q -x ) > gpurun_out/pytest_gemm.log 2>&1
tail -n 6 gpurun_out/pytest_gemm.log
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1
grep -E "agg|whole_step|encoder|proj|dgrad|wgrad" gpurun_out/stage_times.txt
