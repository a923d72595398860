#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for cfg in "1 1" "1 2" "0 1" "0 2"; do
  set -- $cfg
  CLOUDAAE_PIPELINE=$1 CAAE_HPR_CLUSTER=$2 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train_p$1_c$2.json 2> gpurun_out/bench_train_p$1_c$2.err
done
CAAE_HPR_CLUSTER=1 timeout 120 python tools/debug_hpr_timing.py > gpurun_out/hpr_timing_c1.txt 2>&1
timeout 120 python tools/debug_hpr_timing.py > gpurun_out/hpr_timing_c2.txt 2>&1
timeout 300 python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1
tail -n 8 gpurun_out/pytest_gpu.log
cat gpurun_out/hpr_timing_c1.txt gpurun_out/hpr_timing_c2.txt gpurun_out/stage_times.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_train_p*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("stage_ms"), d["e2e"]["value"], d.get("losses_last_step"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
