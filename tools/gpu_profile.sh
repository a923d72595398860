#!/bin/bash
# ncu launch list + full capture of the top kernels (kept under gpurun's 64 MiB return limit).
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/profile_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'hpr_select_kernel|gemm_tf32_kernel|nn_distance_fwd_kernel|fps_reg_kernel|knn_kernel' \
    -c ${NCU_COUNT:-14} -o gpurun_out/prof_r1 python tools/profile_step.py --steps 1 > gpurun_out/prof_full.log 2>&1
du -sh gpurun_out/*
