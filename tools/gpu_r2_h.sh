#!/bin/bash
set -u
mkdir -p gpurun_out
for tc in 1 0; do
  echo "== CAAE_KNN_TC=$tc"
  CAAE_KNN_TC=$tc timeout 600 python -m pytest tests/test_gpu_model.py -q -k "test_train_forward_losses_and_gradients" 2>&1 | grep -E "^E   +Assert|passed|failed" | head -8
done
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | tail -8
