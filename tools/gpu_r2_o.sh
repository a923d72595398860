#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/b128_parity.json
( timeout 900 python -m pytest tests/test_gpu_model_b128.py tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_eval.py -x -q -s ) > gpurun_out/pytest_model.log 2>&1
grep "B=128 tf32 dgcnn" gpurun_out/pytest_model.log | grep -v print | grep -o '"grad_l2_worst[^}]*'; tail -3 gpurun_out/pytest_model.log
timeout 300 python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1; cat gpurun_out/stage_times.txt
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
timeout 300 python bench.py --workload ops --steps 50 2>/dev/null | python -c "
import json,sys; o=json.loads(sys.stdin.readline()); print('ops', o['value'], {k:(round(v['ms']*1e3,1), round(v.get('reference_kernel_ms',0)*1e3,1)) for k,v in o['kernels'].items()})"
CLOUDAAE_BENCH_LIGHT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 4 --warmup 3 > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bench.csv 45 > gpurun_out/launches_bench_summary.txt 2>&1; head -8 gpurun_out/launches_bench_summary.txt
bash tools/sanitize.sh 2>&1 | tail -8
