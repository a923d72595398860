#!/bin/bash
# compute-sanitizer recipe (run on the GPU box): memcheck + racecheck over the kernels with the most intricate
# shared-memory protocols — hidden point removal (work queues, 8-lane groups), EdgeConv backward (shared-memory atomics),
# the tensor-core kNN screen (tcgen05 / TMEM / mbarrier) and the chamfer backward (warp-aggregated scatter).
#   bash tools/sanitize.sh [seconds per selection, default 240]
#   -> gpurun_out/sanitize_{memcheck,racecheck}.log + a one-line verdict each
set -u
LIMIT=${1:-240}
mkdir -p gpurun_out
TESTS=(   # file|-k expression
  "tests/test_gpu_synthesis.py|test_hidden_point_removal_matches_qhull or test_hpr_duplicates_and_padding_draws"
  "tests/test_gpu_model.py|knn_tensor_core_screen_is_bit_identical_to_the_ffma_kernel and 17-8-8-10"
  "tests/test_gpu_model.py|test_knn_matches_reference_topk"
  "tests/test_gpu_ops.py|nn_distance_grad"
  "tests/test_gpu_model.py|test_train_forward_losses_and_gradients and dgcnn-3-128-fp32"
)
for tool in racecheck memcheck; do
  log=gpurun_out/sanitize_$tool.log
  : > $log
  for t in "${TESTS[@]}"; do
    echo "=== $tool: pytest ${t%%|*} -k '${t#*|}'" >> $log
    timeout $LIMIT compute-sanitizer --tool $tool --print-limit 5 --error-exitcode 9 python -m pytest "${t%%|*}" -k "${t#*|}" -x -q -p no:cacheprovider >> $log 2>&1
    echo "=== exit $?" >> $log
  done
  echo "$tool: $(grep -c '=== exit 0' $log) of $(grep -c "^=== $tool:" $log) test selections clean (exit 124 = the ${LIMIT}-s limit); summary lines:"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | sort | uniq -c
done
