#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_evaluation.py tests/test_gpu_eval.py -q -x ) > gpurun_out/pytest_eval2.log 2>&1
tail -n 8 gpurun_out/pytest_eval2.log
timeout 200 python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
tail -n 3 gpurun_out/bench_infer.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_infer.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["front_end"]["ms"], d["icp"])
PY
