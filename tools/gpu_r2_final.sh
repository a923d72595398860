#!/bin/bash
# Round-2 final evidence at HEAD on one B200, most important first, every command under its own timeout.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/b128_parity.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --workload ops --steps 50 > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
timeout 300 python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
def line(p): return json.loads(open(p).read().strip().splitlines()[-1])
try:
    d=line("gpurun_out/bench_train.json")
    print("train ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "step frac", d["roofline"]["step"]["frac"], "cpu", d["cpu_baseline"]["value"])
    for e in d["roofline"]["kernels"]: print("  %-70s %8.1f us  %-8s frac %s" % (e["kernel"][:70], e["us"], e["bound"], None if e["frac"] is None else round(e["frac"],3)))
except Exception as e: print("train line:", e)
try:
    o=line("gpurun_out/bench_ops.json"); print("ops", o["value"], {k:(round(v["ms"]*1e3,1), round(v.get("reference_kernel_ms",0)*1e3,1)) for k,v in o["kernels"].items()})
    i=line("gpurun_out/bench_infer.json"); print("infer", i["value"], i["e2e"]["value"])
    r=line("gpurun_out/bench_reference.json"); print("reference arm", r["value"], r["config"].get("reference_sample_per_step"))
except Exception as e: print("other lines:", e)
PY
# launch list of the bench command (graph kernel nodes are listed individually)
CLOUDAAE_BENCH_LIGHT=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bench.csv 40 > gpurun_out/launches_bench_summary.txt 2>&1; head -8 gpurun_out/launches_bench_summary.txt
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1; head -10 gpurun_out/stage_times.txt
timeout 120 python tools/time_knn.py > gpurun_out/time_knn.txt 2>&1; cat gpurun_out/time_knn.txt
du -sh gpurun_out
