#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (both arms), ncu launch list.  Full ncu captures: tools/gpu_profile.sh.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 300 python bench.py --workload ops > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/profile_step.log 2>&1
du -sh gpurun_out/* | tail -20
tail -3 gpurun_out/pytest_gpu.log
head -c 1500 gpurun_out/bench_train.json
