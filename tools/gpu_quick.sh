#!/bin/bash
# short, hang-proof check of the EdgeConv recording path: every command under its own small timeout
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_model.py tests/test_gpu_model_b128.py tests/test_gpu_eval.py -x -q 2>&1 | tail -2
for r in 0 1 0 1; do echo "EDGE_REC=$r $(CLOUDAAE_EDGE_REC=$r timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1 | cut -c1-60)"; done
timeout 200 python tools/stage_times.py --detail 2>&1 | grep -E "^L4|^L2|encoder|whole" 
