#!/bin/bash
# train lines at 2 and 4 GPUs of one box (run with gpurun --gpus 4)
set -u
mkdir -p gpurun_out
for N in 2 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 150 --warmup 10 2> gpurun_out/bench_train_dp$N.err | grep '^{"metric"' > gpurun_out/bench_train_dp$N.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_train_dp$N.json").read().strip().splitlines()[-1])
    print("train N=$N ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("repeat_ms_per_step"))
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/bench_train_dp$N.err").read()[-800:])
PY
done
timeout 100 python bench.py --steps 150 --warmup 10 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_train_dp1_samebox.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_train_dp1_samebox.json').read().strip().splitlines()[-1]); print('train N=1 (same box) ms/step', round(d['ms_per_step'],4), d.get('repeat_ms_per_step'))"
