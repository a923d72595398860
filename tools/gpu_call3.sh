#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for c in 1 4; do
  CAAE_HPR_CLUSTER=$c timeout 300 python -m pytest tests/test_gpu_synthesis.py -m gpu -x -q > gpurun_out/pytest_synth_c$c.log 2>&1
  echo "cluster $c exit $?" >> gpurun_out/pytest_synth_c$c.log
done
for c in 1 2 4; do
  CAAE_HPR_CLUSTER=$c timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train_c$c.json 2> gpurun_out/bench_train_c$c.err
done
timeout 120 python tools/debug_hpr_timing.py > gpurun_out/hpr_timing.txt 2>&1
timeout 300 python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_synth_c1.log gpurun_out/pytest_synth_c4.log
cat gpurun_out/hpr_timing.txt gpurun_out/stage_times.txt
python - <<'PY'
import json
for f in ("bench_train_c1","bench_train_c2","bench_train_c4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("stage_ms"), d.get("losses_last_step"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
