#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/b128_parity.json
( timeout 900 python -m pytest tests/test_gpu_model_b128.py tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_eval.py -x -q -s ) > gpurun_out/pytest_model.log 2>&1
grep "B=128 tf32 dgcnn" gpurun_out/pytest_model.log | grep -v print | grep -o '"grad_l2_worst[^}]*'; tail -3 gpurun_out/pytest_model.log
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1; grep -E "fwd|bwd|adam|whole|sum|synth" gpurun_out/stage_times.txt
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
