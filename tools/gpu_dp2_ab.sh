#!/bin/bash
# A/B of the NCCL side-stream priority at 2 GPUs.
set -u
mkdir -p gpurun_out
for prio in 0 -2; do
  CLOUDAAE_NCCL_PRIORITY=$prio timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((prio+5)) bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_dp2_prio$prio.json 2> gpurun_out/bench_dp2_prio$prio.err
  echo "prio $prio exit $?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dp2_prio$prio.json").read().strip().splitlines()[-1])
print("prio $prio", d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
done
