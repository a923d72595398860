#!/bin/bash
# last check of the round at HEAD: GPU suite + smoke, then ncu --set full of the step's top kernels reduced on the box
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 ) > gpurun_out/pytest_gpu_last.txt; cat gpurun_out/pytest_gpu_last.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 240 ncu --set full --clock-control none \
    -k regex:'hpr_select_kernel|gemm_tf32_persist|gemm_tf32_big|knn_tc_kernel|nn_distance_fwd_kernel|nn_distance_bwd|edge_cloud_kernel|bn_act_bwd_vec4|bn_act_meanpool|adam_tf' \
    -c 40 -o /tmp/prof_step python tools/profile_step.py --steps 1 > gpurun_out/prof_full.log 2>&1
python tools/ncu_traffic.py /tmp/prof_step.ncu-rep gpurun_out/ncu_traffic_step.json > gpurun_out/ncu_traffic_step.txt 2>&1
python tools/ncu_metrics.py /tmp/prof_step.ncu-rep > gpurun_out/ncu_metrics_step.txt 2>&1
head -12 gpurun_out/ncu_traffic_step.txt; wc -l gpurun_out/ncu_metrics_step.txt
