#!/bin/bash
# round 2, call A: B=128 TF32 parity (observed errors), TF32 peak, stage times at the start of the round
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests/test_gpu_model_b128.py -x -q -s ) > gpurun_out/pytest_b128.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_b128.log
tail -12 gpurun_out/pytest_b128.log
python tools/measure_tf32_peak.py 2 > gpurun_out/tf32_peak.json 2>&1; cat gpurun_out/tf32_peak.json
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times_r2_start.txt 2>&1; tail -50 gpurun_out/stage_times_r2_start.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
