#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1 | tee gpurun_out/step_time_head.txt
