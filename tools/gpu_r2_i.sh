#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_evaluation.py tests/test_gpu_ops.py -x -q ) > gpurun_out/pytest_eval.log 2>&1; tail -5 gpurun_out/pytest_eval.log
