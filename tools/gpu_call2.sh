#!/bin/bash
# GPU parity tests, HPR phase split, train bench with and without stream concurrency.
set -u
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python tools/debug_hpr_timing.py > gpurun_out/hpr_timing.txt 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
CLOUDAAE_STREAMS=0 timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train_nostreams.json 2> gpurun_out/bench_train_nostreams.err
tail -15 gpurun_out/pytest_gpu.log
cat gpurun_out/hpr_timing.txt
python - <<'PY'
import json
for f in ("bench_train","bench_train_nostreams"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("stage_ms"), d.get("e2e"), d.get("losses_last_step"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
