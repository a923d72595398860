#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 300 python bench.py --workload ops --steps 50 --warmup 5 > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1
head -10 gpurun_out/stage_times.txt
cat gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
head -c 1200 gpurun_out/bench_ops.json; echo; head -c 600 gpurun_out/bench_reference.json
