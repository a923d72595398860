#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "knn" ) > gpurun_out/pytest_knn.log 2>&1; tail -3 gpurun_out/pytest_knn.log
python tools/knn_real_once.py 2>&1 | tail -1
timeout 120 python tools/time_knn.py 2>&1 | tee gpurun_out/time_knn.txt
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
