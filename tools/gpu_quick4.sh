#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_model.py tests/test_gpu_evaluation.py -x -q 2>&1 | tail -2
for q in 4 2 1; do CAAE_NND_Q=$q timeout 60 python tools/time_nnd.py 2>&1 | tail -1; done
for q in 4 2; do echo "Q=$q $(CAAE_NND_Q=$q timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1)"; done
