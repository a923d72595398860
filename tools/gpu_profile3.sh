#!/bin/bash
# Round-end evidence, part 2: ncu --set full captures, reduced on the box to CSV / JSON summaries (the .ncu-rep
# files exceed what gpurun copies back).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none \
    -k regex:'hpr_select_kernel|gemm_tf32|knn_kernel|nn_distance_fwd_kernel|fps_reg_kernel|edge_cloud_kernel|bn_act_bwd_vec4|col_reduce_vec4|adam_tf' \
    -c 44 -o /tmp/prof_step python tools/profile_step.py --steps 1 > gpurun_out/prof_full.log 2>&1
timeout 600 ncu --set full --clock-control none \
    -k regex:'segment_extract|radius_count|radius_compact|fps_seeded|icp_refine' --launch-skip 7 -c 7 \
    -o /tmp/prof_eval python tools/eval_once.py > gpurun_out/prof_eval.log 2>&1
python tools/ncu_traffic.py /tmp/prof_step.ncu-rep gpurun_out/ncu_traffic_step.json > gpurun_out/ncu_traffic_step.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_eval.ncu-rep gpurun_out/ncu_traffic_eval.json > gpurun_out/ncu_traffic_eval.txt 2>&1
for n in step eval; do
  python tools/ncu_metrics.py /tmp/prof_$n.ncu-rep > gpurun_out/ncu_metrics_$n.txt 2>&1
done
cat gpurun_out/ncu_traffic_eval.txt; tail -2 gpurun_out/prof_eval.log; du -sh gpurun_out
