"""A/B of the step pipelines (single graph vs two decoupled graphs) at B=128: ms per step over 200 replays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
from cloudaae_b200.train import CloudAAETrainer
dev = torch.device("cuda", 0)
B = bench.TRAIN_B
pool = [{k: torch.from_numpy(v).to(dev) for k, v in bt.items()} for bt in bench.pose_batches(B, seed=0)]
modes = sys.argv[1:] or ["1", "2", "3"]
for mode in modes:
    tr = CloudAAETrainer(batch_size=B, num_point=bench.TRAIN_N, device=dev, seed=0)
    syn = SegmentSynthesizer(load_models_xyz(device=dev), B, bench.TRAIN_N, seed=1234)
    first = [pool[0][k] for k in bench.TRAIN_KEYS]
    if mode == "1":
        static = tr.capture_online_pipelined(syn, *first)
    else:
        static = tr.capture_online_decoupled(syn, *first, depth=int(mode))
    def load(i):
        for dst, k in zip(static, bench.TRAIN_KEYS):
            dst.copy_(pool[i % len(pool)][k], non_blocking=True)
    for i in range(10):
        load(i); tr.replay()
    tr.join(); torch.cuda.synchronize()
    for rep in range(2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for i in range(200):
            load(i); tr.replay()
        host = time.perf_counter() - t0
        tr.join(); e.record(); torch.cuda.synchronize()
        print(f"mode {mode}: {s.elapsed_time(e) / 200:.4f} ms/step  (host enqueue {host / 200 * 1e3:.4f} ms/step) losses {tr.losses.tolist()}", flush=True)
    del tr, syn
