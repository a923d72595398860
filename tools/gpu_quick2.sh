#!/bin/bash
# short, hang-proof A/B of the thin weight-gradient routing: every command under its own small timeout
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gemm_tf32.py -x -q -k "layouts and 24-128-32768" 2>&1 | tail -2
CLOUDAAE_WGRAD_FLOOR=26 timeout 100 python -m pytest tests/test_gpu_model_b128.py -x -q 2>&1 | tail -2
for f in 28 26 28 26; do
  echo "floor $f: $(CLOUDAAE_WGRAD_FLOOR=$f timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1)"
done
