#!/bin/bash
# New evaluation-side kernels: parity tests, the infer workload, a memcheck pass over the new tests.
set -u
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_evaluation.py -q -x ) > gpurun_out/pytest_eval.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_eval.log
tail -n 40 gpurun_out/pytest_eval.log
timeout 300 python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
tail -n 5 gpurun_out/bench_infer.err; cat gpurun_out/bench_infer.json
( timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_evaluation.py -q -x -k "extract or radius or fps_random or one_call" ) > gpurun_out/memcheck_eval.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck_eval.log
tail -n 15 gpurun_out/memcheck_eval.log
