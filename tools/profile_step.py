"""Profiling driver (run under ncu on the GPU box): a few EAGER iterations of the full train step with
on-line synthesis (one C-ABI launch per kernel, no CUDA graph) followed by the tf_ops microbench pass.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --steps 2

Prints the number of C-ABI launches per step so `-s/-c` windows can be chosen.  Never a bench value.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import cloudaae_b200 as caae  # noqa: E402
from cloudaae_b200 import _capi  # noqa: E402
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz  # noqa: E402
from cloudaae_b200.train import CloudAAETrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--what", default="train,ops")
    ap.add_argument("--batch", type=int, default=bench.TRAIN_B)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    what = args.what.split(",")
    if "ops" in what:
        B = bench.OPS_B
        clouds_h, pred_h, target_h = bench.ops_inputs(B, seed=0)
        clouds, pred, target = (torch.from_numpy(a).to(dev) for a in (clouds_h, pred_h, target_h))
        g = torch.full((B, bench.OPS_CH), 1.0 / (B * bench.OPS_CH), device=dev)
        for i in range(args.steps):
            caae.farthest_point_sample_gather(bench.OPS_M, clouds)
            _, i1, _, i2 = caae.nn_distance(pred, target)
            caae.nn_distance_grad(pred, target, g, i1, g, i2)
            torch.cuda.synchronize()
        print("ops pass done", flush=True)

    if "train" in what:
        B = args.batch
        tr = CloudAAETrainer(batch_size=B, num_point=bench.TRAIN_N, device=dev, seed=0)
        syn = SegmentSynthesizer(load_models_xyz(device=dev), B, bench.TRAIN_N, seed=1234)
        pool = bench.pose_batches(B, seed=0, pool=2)
        pool_d = [{k: torch.from_numpy(v).to(dev) for k, v in bt.items()} for bt in pool]
        for i in range(args.steps):
            before = _capi.COUNTER[0]
            bt = pool_d[i % len(pool_d)]
            tr.train_step_online(syn, *[bt[k] for k in bench.TRAIN_KEYS])
            torch.cuda.synchronize()
            print(f"train step {i}: {_capi.COUNTER[0] - before} C-ABI launches, losses {tr.losses.tolist()}", flush=True)


if __name__ == "__main__":
    main()
