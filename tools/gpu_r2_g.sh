#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "knn" ) > gpurun_out/pytest_knn.log 2>&1; tail -3 gpurun_out/pytest_knn.log
timeout 120 python tools/time_knn.py 2>&1 | tee gpurun_out/time_knn.txt
rm -f gpurun_out/b128_parity.json
( timeout 900 python -m pytest tests/test_gpu_model_b128.py -q -s ) > gpurun_out/pytest_b128.log 2>&1
grep "B=128" gpurun_out/pytest_b128.log | grep -v print; tail -3 gpurun_out/pytest_b128.log
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1; cat gpurun_out/stage_times.txt
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
