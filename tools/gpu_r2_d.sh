#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "knn" ) > gpurun_out/pytest_knn.log 2>&1; tail -15 gpurun_out/pytest_knn.log
timeout 120 python tools/time_knn.py 2>&1 | tee gpurun_out/time_knn.txt
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_model_b128.py tests/test_gpu_eval.py -x -q ) > gpurun_out/pytest_model.log 2>&1; tail -3 gpurun_out/pytest_model.log
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
