#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_dp8.json 2> gpurun_out/bench_dp8.err
echo "exit $?"
tail -n 4 gpurun_out/bench_dp8.err | cut -c1-400
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dp8.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"])
PY
