#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 300 python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/profile_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'hpr_select_kernel|knn_kernel' -c 6 -o gpurun_out/prof_r1c python tools/profile_step.py --steps 1 --what train > gpurun_out/prof_full.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log
cat gpurun_out/stage_times.txt
head -c 900 gpurun_out/bench_train.json
du -sh gpurun_out/prof_r1c.ncu-rep
