#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gemm_tf32.py -x -q ) > gpurun_out/pytest_gemm.log 2>&1; tail -4 gpurun_out/pytest_gemm.log
rm -f gpurun_out/b128_parity.json
( timeout 900 python -m pytest tests/test_gpu_model_b128.py tests/test_gpu_model.py tests/test_gpu_eval.py -x -q -s ) > gpurun_out/pytest_model.log 2>&1
grep "B=128 tf32 eval" gpurun_out/pytest_model.log | head -1; tail -3 gpurun_out/pytest_model.log
timeout 300 python bench.py --workload infer --steps 3 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; tail -2 gpurun_out/bench_infer.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_infer.json").read().strip().splitlines()[0])
print("infer", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
PY
