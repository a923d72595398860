#!/bin/bash
# 2-GPU: correctness smoke (parameters stay identical across ranks), then the train bench with NCCL CTA caps
set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 tools/dp_smoke.py 2>&1 | grep -E "params|done|Error|error" | head -12
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_dp2_$name.json 2> gpurun_out/bench_dp2_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_dp2_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(d["ms_per_step"],4), "seg/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_dp2_$name.err").read()[-800:])
PY
}
timeout 300 python bench.py --steps 200 --warmup 10 --workload train > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print('N=1 ms/step', round(d['ms_per_step'],4), 'seg/s', round(d['value']))"
run default A=1
run maxctas8 NCCL_MAX_CTAS=8
run maxctas4 NCCL_MAX_CTAS=4
run maxctas2 NCCL_MAX_CTAS=2
