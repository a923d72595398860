#!/bin/bash
# usage: bash tools/gpu_r2_scale.sh N   — train and inference lines at N GPUs (one box), both arms of the train workload
set -u
N=$1
mkdir -p gpurun_out
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,TUNING
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_train_dp$N.json 2> gpurun_out/bench_train_dp$N.err
unset NCCL_DEBUG NCCL_DEBUG_SUBSYS
grep -E "NCCL INFO (Connected|Channel 00|comm .* nranks|.*NVLS|.*algo|Using network|Trees|[0-9]+ coll channels)" gpurun_out/bench_train_dp$N.err | head -12 > gpurun_out/nccl_dp$N.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --workload infer --steps 5 --warmup 3 > gpurun_out/bench_infer_dp$N.json 2> gpurun_out/bench_infer_dp$N.err
python - <<PY
import json
for w in ("train", "infer"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{w}_dp$N.json").read().strip().splitlines()[-1])
        print(w, "N=$N ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("clocks"))
    except Exception as e:
        print(w, "failed", e); print(open(f"gpurun_out/bench_{w}_dp$N.err").read()[-600:])
PY
cat gpurun_out/nccl_dp$N.txt | cut -c1-200 | head -8
