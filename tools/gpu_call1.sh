#!/bin/bash
# Session re-entry check: GPU parity tests, HPR phase split, train bench, full ncu capture of the top kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python tools/debug_hpr_timing.py > gpurun_out/hpr_timing.txt 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'hpr_select_kernel|gemm_tf32_kernel|nn_distance_fwd_kernel|fps_reg_kernel|knn_kernel|edge_cloud_kernel|gemm_simt_kernel' \
    -c 40 -o gpurun_out/prof_r1b python tools/profile_step.py --steps 1 > gpurun_out/prof_full.log 2>&1
du -sh gpurun_out/* | tail -20
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/hpr_timing.txt
head -c 600 gpurun_out/bench_train.json
