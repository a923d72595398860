#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
tail -n 3 gpurun_out/bench_train.err
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1
cat gpurun_out/stage_times.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_train.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["stage_ms"], d["e2e"]["value"], d["roofline"]["all_agg_gemms_ms"], d["synthesis_kernel"]["ms_per_launch"], d["cpu_baseline"])
PY
