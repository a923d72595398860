#!/bin/bash
# launch list of the eager step + full ncu capture of the top kernels (source imported for hpr / knn / gemm only via -lineinfo)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/profile_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'hpr_select_kernel|gemm_tf32|knn_kernel|nn_distance_fwd_kernel|fps_reg_kernel|edge_cloud_kernel|bn_act_bwd_vec4|col_reduce_vec4|adam_tf' \
    -c 44 -o gpurun_out/prof_r1d python tools/profile_step.py --steps 1 > gpurun_out/prof_full.log 2>&1
du -sh gpurun_out/prof_r1d.ncu-rep; tail -3 gpurun_out/prof_full.log
