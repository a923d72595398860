#!/bin/bash
# Round-end evidence, part 1: GPU parity tests, smoke, bench lines (both arms, all workloads), ncu launch list.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 300 python bench.py --workload ops > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
timeout 300 python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/profile_step.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_train.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["all_agg_gemms_ms"], d["cpu_baseline"]["value"])
PY
du -sh gpurun_out
