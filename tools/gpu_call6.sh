#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1
timeout 120 python tools/debug_hpr_timing.py > gpurun_out/hpr_timing.txt 2>&1
tail -n 6 gpurun_out/pytest_gpu.log
cat gpurun_out/stage_times.txt gpurun_out/hpr_timing.txt
head -c 300 gpurun_out/bench_train.json
