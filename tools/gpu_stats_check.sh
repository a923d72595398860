#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_gemm_tf32.py -q -x -k "fused or persistent" ) > gpurun_out/pytest_stats.log 2>&1
tail -n 12 gpurun_out/pytest_stats.log
( timeout 200 python -m pytest tests/test_gpu_model.py tests/test_gpu_synthesis.py -q -x ) > gpurun_out/pytest_model.log 2>&1
tail -n 6 gpurun_out/pytest_model.log
timeout 100 python tools/ab_pipeline.py 1 2>&1 | tail -2
CLOUDAAE_FUSED_STATS=0 timeout 100 python tools/ab_pipeline.py 1 2>&1 | tail -1
