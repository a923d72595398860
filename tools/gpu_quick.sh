#!/bin/bash
# short, hang-proof check: every command under its own small timeout
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_model_b128.py tests/test_gpu_eval.py tests/test_gpu_synthesis.py tests/test_gpu_smoke.py -x -q 2>&1 | tail -2
python - <<'PY'
import json
d=json.load(open('gpurun_out/b128_parity.json'))
for k,v in d.items(): print(k, {kk:(round(vv,7) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('embedding','rot','loss_total','grad_l2_worst','grad_l2_worst_name','grad_l2_median')})
PY
for r in 0 1 0 1; do echo "EDGE_REC=$r $(CLOUDAAE_EDGE_REC=$r timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1 | cut -c1-40)"; done
