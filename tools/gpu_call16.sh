#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_synthesis.py -q -x -k "decoupled or pipelined" ) > gpurun_out/pytest_dec.log 2>&1
tail -n 15 gpurun_out/pytest_dec.log
timeout 300 python tools/ab_pipeline.py 1 2 3 2>&1 | tee gpurun_out/ab_pipeline.txt | tail -12
