#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err
echo "exit $?"
tail -c 1500 gpurun_out/bench_dp2.json; tail -5 gpurun_out/bench_dp2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_dp2.json 2> gpurun_out/bench_ref_dp2.err
echo "ref exit $?"; head -c 400 gpurun_out/bench_ref_dp2.json
