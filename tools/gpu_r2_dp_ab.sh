#!/bin/bash
# usage: bash tools/gpu_r2_dp_ab.sh N — data-parallel step at N GPUs: single-graph pipeline vs the 3-slot decoupled
# queue (absorbs the per-batch variance of the synthesis cost, which a synchronous allreduce otherwise turns into
# max-over-ranks every step), each with NCCL's default CTA count and with 8 CTAs
set -u
N=$1
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus $N --steps 150 --warmup 10 2> gpurun_out/dp_ab_$name.err | grep '^{"metric"' > gpurun_out/dp_ab_${name}_dp$N.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dp_ab_${name}_dp$N.json").read().strip().splitlines()[-1])
    print("$name N=$N ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/dp_ab_$name.err").read()[-800:])
PY
}
run base CLOUDAAE_PIPELINE=1
run q3 CLOUDAAE_PIPELINE=3
run base_cta8 CLOUDAAE_PIPELINE=1 CLOUDAAE_NCCL_MAX_CTAS=8
run q3_cta8 CLOUDAAE_PIPELINE=3 CLOUDAAE_NCCL_MAX_CTAS=8
