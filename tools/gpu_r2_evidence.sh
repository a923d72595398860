#!/bin/bash
# Round-2 evidence on one B200: GPU parity tests, smoke, bench lines (both arms, all workloads), ncu launch list of the
# bench command, ncu --set full of the step's top kernels reduced on the box, compute-sanitizer recipe.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/b128_parity.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 400 python bench.py --workload ops --steps 50 > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
timeout 400 python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_train.json").read().strip().splitlines()[0])
print("train ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "step frac", d["roofline"]["step"]["frac"], "cpu", d["cpu_baseline"]["value"])
for e in d["roofline"]["kernels"]: print("  %-70s %8.1f us  %-8s frac %s" % (e["kernel"][:70], e["us"], e["bound"], None if e["frac"] is None else round(e["frac"],3)))
o=json.loads(open("gpurun_out/bench_ops.json").read().strip().splitlines()[0]); print("ops", o["value"], {k:(round(v["ms"]*1e3,1), round(v.get("reference_kernel_ms",0)*1e3,1)) for k,v in o["kernels"].items()})
i=json.loads(open("gpurun_out/bench_infer.json").read().strip().splitlines()[0]); print("infer", i["value"], i["e2e"]["value"])
r=json.loads(open("gpurun_out/bench_reference.json").read().strip().splitlines()[0]); print("reference arm", r["value"], r["config"].get("reference_sample_per_step"))
PY
# launch list of the SAME command as the bench line (graph kernel nodes are listed individually); first 4000 launches
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bench.csv 40 > gpurun_out/launches_bench_summary.txt 2>&1; head -12 gpurun_out/launches_bench_summary.txt
# eager step + ops pass: one launch list that maps 1:1 onto C-ABI calls
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step.csv 45 > gpurun_out/launches_step_summary.txt 2>&1
# ncu --set full of the top kernels, reduced on the box
timeout 1200 ncu --set full --clock-control none \
    -k regex:'hpr_select_kernel|gemm_tf32|knn_tc_kernel|knn_kernel|nn_distance_fwd_kernel|nn_distance_bwd|fps_reg_kernel|edge_cloud_kernel|bn_act_bwd_vec4|bn_act_meanpool|adam_tf' \
    -c 60 -o /tmp/prof_step python tools/profile_step.py --steps 1 > gpurun_out/prof_full.log 2>&1
python tools/ncu_traffic.py /tmp/prof_step.ncu-rep gpurun_out/ncu_traffic_step.json > gpurun_out/ncu_traffic_step.txt 2>&1
python tools/ncu_metrics.py /tmp/prof_step.ncu-rep > gpurun_out/ncu_metrics_step.txt 2>&1
head -25 gpurun_out/ncu_traffic_step.txt
bash tools/sanitize.sh 2>&1 | tail -8
du -sh gpurun_out
