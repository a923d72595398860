#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/b128_parity.json
( timeout 600 python -m pytest tests/test_gpu_gemm_tf32.py -x -q ) > gpurun_out/pytest_gemm.log 2>&1; tail -5 gpurun_out/pytest_gemm.log
timeout 120 python tools/time_gemm_x3.py 2>&1 | tee gpurun_out/time_gemm_x3.txt
( timeout 900 python -m pytest tests/test_gpu_model_b128.py -q -s ) > gpurun_out/pytest_b128.log 2>&1
grep "B=128" gpurun_out/pytest_b128.log | grep -v print; tail -3 gpurun_out/pytest_b128.log
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_eval.py -x -q ) > gpurun_out/pytest_model.log 2>&1; tail -3 gpurun_out/pytest_model.log
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
