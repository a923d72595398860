#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list and full captures of the top kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 300 python bench.py --workload ops > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/profile_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'hpr_select_kernel|gemm_tf32_kernel|nn_distance_fwd_kernel|fps_reg_kernel|knn_kernel|edge_cloud_kernel|synth_points_kernel|nn_distance_bwd' \
    -s 8 -c 45 -o gpurun_out/prof_r1 python tools/profile_step.py --steps 2 > gpurun_out/prof_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_train.json | head -c 3000
