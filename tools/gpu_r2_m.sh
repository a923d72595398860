#!/bin/bash
set -u
mkdir -p gpurun_out
for pdl in 1 0; do
  echo "== CAAE_PDL=$pdl"
  CAAE_PDL=$pdl timeout 120 python tools/ab_pipeline.py 1 2>&1 | tail -1
  CAAE_PDL=$pdl timeout 300 python tools/stage_times.py 2>&1 | grep -E "encoder|fc_|losses|adam|whole|sum"
done
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
