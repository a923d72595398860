#!/bin/bash
# short, hang-proof check: every command under its own small timeout
set -u
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -2
timeout 100 python -m pytest tests/test_gpu_model_b128.py -x -q 2>&1 | tail -2
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
timeout 100 python bench.py --workload ops --steps 50 2>/dev/null | python -c "
import json,sys; o=json.loads(sys.stdin.readline()); print('ops', o['value'], {k:(round(v['ms']*1e3,1), round(v.get('reference_kernel_ms',0)*1e3,1)) for k,v in o['kernels'].items()})"
