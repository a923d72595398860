#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "knn" ) > gpurun_out/pytest_knn.log 2>&1; tail -3 gpurun_out/pytest_knn.log
timeout 120 python tools/time_knn.py 2>&1 | tee gpurun_out/time_knn.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -3 gpurun_out/bench_train.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_train.json").read().strip().splitlines()[0])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "step frac", d["roofline"]["step"]["frac"])
for e in d["roofline"]["kernels"]: print("  %-75s %8.1f us  %-8s frac %s" % (e["kernel"][:75], e["us"], e["bound"], None if e["frac"] is None else round(e["frac"],3)))
print(d["stage_ms"])
print({k: (v["ms"], v.get("reference_kernel_ms")) for k, v in d["ops_microbench"]["kernels"].items()})
print("infer", d["infer"]["value"], d["infer"]["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
PY
