#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm_tf32.py -m gpu -x -q ) > gpurun_out/pytest_gemm.log 2>&1
echo "gemm pytest exit $?" >> gpurun_out/pytest_gemm.log
tail -n 12 gpurun_out/pytest_gemm.log
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
CAAE_GEMM_BIG=0 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_train_nobig.json 2> gpurun_out/bench_train_nobig.err
timeout 300 python tools/stage_times.py --detail > gpurun_out/stage_times.txt 2>&1
tail -n 6 gpurun_out/pytest_gpu.log
grep -E "encoder|whole|agg|L4 proj" gpurun_out/stage_times.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_train.json")+glob.glob("gpurun_out/bench_train_nobig.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"]["all_agg_gemms_ms"], d.get("losses_last_step"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-800:])
PY
