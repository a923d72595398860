#!/usr/bin/env python
"""bench.py — CloudAAE hot-path benchmark (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|ops|infer]

One rank per GPU (torchrun for N > 1).  W untimed warm-up steps, then EXACTLY K timed steps
bracketed by barrier + synchronize; rank 0 prints ONE JSON line.  Device time comes from CUDA
events on the launching stream; the L2 is flushed between timed iterations (outside the event
pairs).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# ----------------------------------------------------------------------------------------------
def ncu_traffic(kernel_prefix: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed ncu --set full capture
    (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py; the round-1 file as a fallback); (None, reason) when
    absent.  Of several variants with the prefix (template arguments, grids) the longest-running one is taken."""
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            data = json.load(open(path))
        except (OSError, ValueError):
            continue
        hits = [(k["duration_us"], kn, k) for kn, k in data.get("kernels", {}).items() if kn.startswith(kernel_prefix)]
        if hits:
            _, kn, k = max(hits)
            return k["dram_bytes"], f"profiles/{name} ({data.get('source')}): {kn}, {k['duration_us']:.1f} us under ncu"
    return None, f"{kernel_prefix}: no committed ncu capture"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops", 0)),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", 0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measure_tf32_peak(seconds: float = 1.0):
    """Dense TF32 peak of THIS GPU, measured like MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3, allow_tf32):
    burst = best of 10 (the denominator for a kernel timed in isolation), sustained = back to back for `seconds`."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import measure_tf32_peak as M
    return M.measure(seconds)


FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 148 SMs x 128 lanes x 2 FLOP x 1.965 GHz = 74.4


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# Workload "ops": BASELINE.json configs[1] — the tf_ops microbench, batch 32:
#   FPS 2048 -> 256 (+ fused gather) on 32 YCB-shaped clouds, then chamfer nn_distance forward and
#   backward on 1024 x 1024 points (decoder-shaped prediction vs the first 1024 points of the cloud).
# A "segment" is one cloud taken through that pass.
OPS_B, OPS_N, OPS_M, OPS_CH = 32, 2048, 256, 1024


def ycb_shaped_clouds(b: int, seed: int) -> np.ndarray:
    """b posed YCB models, f32[b,2048,3]: committed fixture models x fixture poses (synthetic
    selection by seed).  No occluder / visibility here — that belongs to the train workload."""
    models = np.load(os.path.join(ROOT, "tests", "golden", "ycb_models_xyz.npy"))
    z = np.load(os.path.join(ROOT, "tests", "golden", "ycb_poses.npz"))
    rng = np.random.default_rng(seed)
    sel = rng.integers(0, len(z["class_id"]), b)
    cls = z["class_id"][sel]
    ax = z["axisangle"][sel].astype(np.float64)
    theta = np.linalg.norm(ax, axis=1, keepdims=True)
    k = ax / np.maximum(theta, 1e-12)
    K = np.zeros((b, 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    R = np.eye(3)[None] + np.sin(theta)[:, :, None] * K + (1 - np.cos(theta))[:, :, None] * (K @ K)
    pts = models[cls] @ np.transpose(R, (0, 2, 1)).astype(np.float32) + z["translation"][sel][:, None, :]
    return np.ascontiguousarray(pts, dtype=np.float32)


def ops_inputs(b: int, seed: int):
    clouds = ycb_shaped_clouds(b, seed)
    rng = np.random.default_rng(seed + 1000)
    target = clouds[:, :OPS_CH, :]
    pred = (target[:, rng.permutation(OPS_CH), :] + rng.standard_normal((b, OPS_CH, 3)).astype(np.float32) * 0.01)
    return clouds, np.ascontiguousarray(pred, np.float32), np.ascontiguousarray(target, np.float32)


def reference_gpu_kernel_times(clouds, pred, target, gscale, i1, i2, time_kernel, iters):
    """ms per launch of the reference's CUDA launchers (farthestpointsamplingLauncher tf_sampling_g.cu:203-205,
    NmDistanceKernelLauncher :128-131, NmDistanceGradKernelLauncher :152-157) compiled for sm_100a from /root/reference
    into oracle/_ref; None when that library was not built.  They launch on the legacy default stream = torch's
    default stream, so the same CUDA events bracket them."""
    import ctypes

    import torch
    try:
        from oracle import ops as O
        if not O.have_ref():
            return None
        R = O.ref()
    except Exception:
        return None
    b, n, _ = clouds.shape
    m = pred.shape[1]
    dp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    ci = ctypes.c_int
    temp = torch.empty(32, n, device=clouds.device)
    out = torch.empty(b, OPS_M, dtype=torch.int32, device=clouds.device)
    d1 = torch.empty(b, m, device=clouds.device); d2 = torch.empty(b, m, device=clouds.device)
    j1 = torch.empty(b, m, dtype=torch.int32, device=clouds.device); j2 = torch.empty_like(j1)
    g1 = torch.empty(b, m, 3, device=clouds.device); g2 = torch.empty(b, m, 3, device=clouds.device)
    t_fps = time_kernel(lambda: R.ref_gpu_fps(ci(b), ci(n), ci(OPS_M), dp(clouds), dp(temp), dp(out)), iters)
    t_fwd = time_kernel(lambda: R.ref_gpu_nn_distance(ci(b), ci(m), dp(pred), ci(m), dp(target), dp(d1), dp(j1), dp(d2), dp(j2)), iters)
    t_bwd = time_kernel(lambda: R.ref_gpu_nn_distance_grad(ci(b), ci(m), dp(pred), ci(m), dp(target), dp(gscale), dp(i1), dp(gscale),
                                                           dp(i2), dp(g1), dp(g2)), iters)
    return {"fps": t_fps, "nn_distance_fwd": t_fwd, "nn_distance_bwd": t_bwd,
            "what": "the reference's .cu files compiled unmodified with nvcc -arch=sm_100a (oracle/build_ref.sh), same inputs"}


def run_ours_ops(args, rank, world, local_rank, with_cpu_baseline=True):
    import torch
    import torch.distributed as dist

    import cloudaae_b200 as caae

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = measured_peaks()
    B = OPS_B
    clouds_h, pred_h, target_h = ops_inputs(B, seed=rank)
    clouds = torch.from_numpy(clouds_h).to(dev)
    pred = torch.from_numpy(pred_h).to(dev)
    target = torch.from_numpy(target_h).to(dev)
    gscale = torch.full((B, OPS_CH), 1.0 / (B * OPS_CH), device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2
    stream = torch.cuda.current_stream()

    def step():
        idx, sub = caae.farthest_point_sample_gather(OPS_M, clouds)
        d1, i1, d2, i2 = caae.nn_distance(pred, target)
        g1, g2 = caae.nn_distance_grad(pred, target, gscale, i1, gscale, i2)
        return idx, sub, d1, d2, g1

    # --- per-kernel timing for the roofline block (CUDA events on the launching stream)
    def time_kernel(fn, iters):
        evs = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream); fn(); e.record(stream)
            evs.append((s, e))
        torch.cuda.synchronize()
        return float(np.mean([s.elapsed_time(e) for s, e in evs]))  # ms

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    events = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        step()
        e.record(stream)
        events.append((s, e))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    total_ms = float(sum(s.elapsed_time(e) for s, e in events))
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # --- per-kernel breakdown + roofline of the dominant kernel
    k_iters = max(5, min(args.steps, 50))
    t_fps = time_kernel(lambda: caae.farthest_point_sample_gather(OPS_M, clouds), k_iters)
    _, i1, _, i2 = caae.nn_distance(pred, target)
    t_nnd = time_kernel(lambda: caae.nn_distance(pred, target), k_iters)
    t_bwd = time_kernel(lambda: caae.nn_distance_grad(pred, target, gscale, i1, gscale, i2), k_iters)
    fps_bytes = (12 * OPS_N + 4 * OPS_M) * B                 # 12n + 4m per cloud
    nnd_bytes = 20 * (OPS_CH + OPS_CH) * B                   # 20(n+m) per cloud pair
    bwd_bytes = 32 * (OPS_CH + OPS_CH) * B                   # 32(n+m) per cloud pair
    kernels = {
        "fps_2048_256": {"ms": t_fps, "algorithmic_bytes": fps_bytes, "gbs": fps_bytes / t_fps / 1e6,
                         "rounds_per_s_per_cloud": (OPS_M - 1) / (t_fps * 1e-3),
                         "gflops": 8.0 * OPS_N * (OPS_M - 1) * B / t_fps / 1e6},
        "nn_distance_fwd_1024": {"ms": t_nnd, "algorithmic_bytes": nnd_bytes, "gbs": nnd_bytes / t_nnd / 1e6,
                                 "gpairs_per_s": 2.0 * OPS_CH * OPS_CH * B / t_nnd / 1e6,
                                 "gflops": 16.0 * OPS_CH * OPS_CH * B / t_nnd / 1e6},
        "nn_distance_bwd_1024": {"ms": t_bwd, "algorithmic_bytes": bwd_bytes, "gbs": bwd_bytes / t_bwd / 1e6},
    }
    # BASELINE.md section 2 — the kernel-level bar: the reference's OWN CUDA kernels (tf_sampling_g.cu, tf_nndistance_g.cu)
    # rebuilt for sm_100a by oracle/build_ref.sh, same box, same inputs, same event timing.  Comparator only.
    ref_kernels = None
    if rank == 0:
        ref_kernels = reference_gpu_kernel_times(clouds, pred, target, gscale, i1, i2, time_kernel, k_iters)
        if ref_kernels:
            for ours, theirs in (("fps_2048_256", "fps"), ("nn_distance_fwd_1024", "nn_distance_fwd"), ("nn_distance_bwd_1024", "nn_distance_bwd")):
                kernels[ours]["reference_kernel_ms"] = ref_kernels[theirs]
                kernels[ours]["speedup_vs_reference_kernel"] = ref_kernels[theirs] / kernels[ours]["ms"]
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": kernels[dom]["gbs"] / peaks["hbm_gbs"], "traffic": None,
                "peak_source": peaks["source"],
                "note": "FPS is a 255-round dependent chain (latency bound); GB/s on compulsory bytes as the metric asks"}

    # --- e2e: the same pass through the public API with HOST buffers (pinned), copies inside the timed region
    ph_c = torch.from_numpy(clouds_h).pin_memory(); ph_p = torch.from_numpy(pred_h).pin_memory()
    ph_t = torch.from_numpy(target_h).pin_memory()
    out_idx = torch.empty((B, OPS_M), dtype=torch.int32).pin_memory()
    out_loss = torch.empty((B, OPS_CH), dtype=torch.float32).pin_memory()
    out_g = torch.empty((B, OPS_CH, 3), dtype=torch.float32).pin_memory()

    def e2e_step():
        c = ph_c.to(dev, non_blocking=True); p = ph_p.to(dev, non_blocking=True); t = ph_t.to(dev, non_blocking=True)
        idx, sub = caae.farthest_point_sample_gather(OPS_M, c)
        d1, i1, d2, i2 = caae.nn_distance(p, t)
        g1, _ = caae.nn_distance_grad(p, t, gscale, i1, gscale, i2)
        out_idx.copy_(idx, non_blocking=True); out_loss.copy_(d1 + d2, non_blocking=True)
        out_g.copy_(g1, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = ph_c.numel() * 4 + ph_p.numel() * 4 + ph_t.numel() * 4
    d2h = out_idx.numel() * 4 + out_loss.numel() * 4 + out_g.numel() * 4

    if rank != 0:
        return None
    ms_per_step = total_ms / args.steps
    result = {
        "metric": "segments/sec (tf_ops microbench pass: FPS 2048->256 + gather + nn_distance fwd+bwd 1024x1024)",
        "value": B * world / (ms_per_step * 1e-3), "unit": "segments/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (committed YCB model fixture x fixture poses)",
        "config": {"workload": "tf_ops microbench, BASELINE.json configs[1]", "batch_per_gpu": B, "fps": [OPS_N, OPS_M],
                   "nn_distance": [OPS_CH, OPS_CH], "l2": "flushed between timed iterations (256 MB write)",
                   "parallelism": f"independent clouds sharded over {world} rank(s), no collective"},
        "roofline": roofline, "kernels": kernels, "reference_kernels_sm100a": ref_kernels,
        "e2e": {"value": B * world / (e2e_s / args.steps), "unit": "segments/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": 3 * args.steps, "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    if with_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline_ops(sample_clouds=8)
    return result



# ----------------------------------------------------------------------------------------------
# Workload "train": BASELINE.json configs[2]/[3] — the full train_cloudAAE_ycbv.py step at the repo
# default batch (128 segments per GPU, num_point 256, k 10): ON-LINE SYNTHESIS from pose records
# (pose transform, spherical occluders, hidden point removal x2, visible-prefix selection, sensor
# noise), input prep, get_model_dgcnn_mean_6d forward, chamfer + pose losses, backward,
# (NCCL allreduce of the flat gradient), TF-style Adam.  One CUDA graph per step.
TRAIN_B, TRAIN_N = 128, 256
TRAIN_KEYS = ("class_id", "axisangle", "translation")
TRAIN_METRIC = "train segments/sec"


def pose_batches(b: int, seed: int, pool: int = 8):
    """`pool` batches of b pose records drawn uniformly from the committed fixture poses (host arrays)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ycb_poses.npz"))
    out = []
    for i in range(pool):
        sel = np.random.default_rng(seed * 1000 + i).integers(0, len(z["class_id"]), b)
        out.append({"class_id": z["class_id"][sel].astype(np.int32), "axisangle": z["axisangle"][sel].astype(np.float32),
                    "translation": z["translation"][sel].astype(np.float32)})
    return out


def train_config(world: int, extra=None):
    cfg = {"workload": "train_cloudAAE_ycbv.py step, BASELINE.json configs[2]" + ("/[3]" if world > 1 else ""),
           "global_batch": TRAIN_B * world, "batch_per_gpu": TRAIN_B, "num_point": TRAIN_N, "k_neighbor": 10,
           "network": "get_model_dgcnn_mean_6d", "synthesis": "on-line, inside the timed step",
           "parallelism": f"dp{world}: NCCL allreduce of the flat fp32 gradient, 2 buckets" if world > 1 else "single GPU"}
    if extra:
        cfg.update(extra)
    return cfg


def run_ours_train(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from cloudaae_b200 import _capi
    from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
    from cloudaae_b200.train import CloudAAETrainer

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = measured_peaks()
    B = TRAIN_B
    pg = dist.group.WORLD if world > 1 else None
    tr = CloudAAETrainer(batch_size=B, num_point=TRAIN_N, device=dev, seed=0, process_group=pg)
    syn = SegmentSynthesizer(load_models_xyz(device=dev), B, TRAIN_N, seed=1234 + rank)
    pool_h = pose_batches(B, seed=rank)
    pool_d = [{k: torch.from_numpy(v).to(dev) for k, v in bt.items()} for bt in pool_h]
    # one CUDA graph per step; by default the synthesis of batch i+1 runs as a parallel branch next to the
    # train step of batch i (the reference's tf.data prefetch(1), train_cloudAAE_ycbv.py:115)
    mode = os.environ.get("CLOUDAAE_PIPELINE", "1")
    pipelined = mode != "0"
    capture = {"0": tr.capture_online, "1": tr.capture_online_pipelined}.get(mode, tr.capture_online_decoupled)
    kw = {"depth": int(mode)} if mode in ("2", "3", "4") else {}
    static = capture(syn, *[pool_d[0][k] for k in TRAIN_KEYS], **kw)

    def load(i, src):  # refresh the graph's static pose records
        for dst, k in zip(static, TRAIN_KEYS):
            dst.copy_(src[i % len(src)][k], non_blocking=True)

    for i in range(args.warmup):
        load(i, pool_d); tr.replay()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        load(i, pool_d); tr.replay()
    tr.join()   # decoupled pipeline: the synthesis running ahead on its own stream ends inside the timed region
    e.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    losses = tr.losses.cpu().tolist()
    # run-to-run spread: four more untouched repetitions of the same K steps (this rank's device time; the reported
    # `value` is the first, contract-timed one)
    repeats = [total_ms / args.steps]
    for _ in range(4):
        s.record()
        for i in range(args.steps):
            load(i, pool_d); tr.replay()
        tr.join()
        e.record()
        torch.cuda.synchronize()
        repeats.append(s.elapsed_time(e) / args.steps)

    # ---- e2e: HOST pose records (pinned) -> H2D -> synthesis + step -> D2H of the loss vector, every step
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in bt.items()} for bt in pool_h]
    loss_h = torch.empty(4, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        load(i, pinned)
        tr.replay()
        loss_h.copy_(tr.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(3):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
    if rank != 0:
        return None

    ms = total_ms / args.steps
    hbm = peaks["hbm_gbs"]
    light = os.environ.get("CLOUDAAE_BENCH_LIGHT", "0") == "1"   # ncu launch-list pass: the timed region only
    if world > 1 or light:
        # data-parallel runs: the per-kernel table needs eager single-rank replays of collectives-free stages; it is
        # reported by the N = 1 run of the same commit.  Here: the step against its composite ceiling, per GPU.
        tf32_s = peaks["bf16_tflops_sustained"] / 2.0
        ceiling_ms = 140e9 / (tf32_s * 1e12) * 1e3 + 5.4e9 / (FP32_FMA_PEAK_TFLOPS * 1e12) * 1e3 + 0.55e9 / (hbm * 1e9) * 1e3
        roofline = {"kernel": "train step (per GPU) vs SURVEY 8(d) composite ceiling; per-kernel table: see the N=1 line",
                    "bound": "tensor", "achieved": 140e9 / (ms * 1e-3) / 1e12, "peak": tf32_s, "unit": "TFLOP/s",
                    "frac": ceiling_ms / ms, "traffic": None,
                    "peak_source": peaks["source"] + " (bf16 sustained / 2)",
                    "step": {"achieved_segments_per_s": B * world / (ms * 1e-3), "ceiling_ms": ceiling_ms, "frac": ceiling_ms / ms}}
        synthesis_kernel = None
        stage_ms = {}
    else:
        # ---- per-kernel table and roofline (rank 0, after the timed region): every kernel family of the step is captured
        # as its own CUDA graph and replayed (tools/stage_times.py), so each time is a warm, launch-gap-free device time
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import stage_times
        c, ax, tl = static
        tf32 = measure_tf32_peak(1.0)
        rows = {name: (ms_, n_) for name, ms_, n_ in stage_times.measure(tr, syn, c, ax, tl, iters=20, detail=True)}
        R, k = B * TRAIN_N, 10
        n_all = syn.nm + syn.no

        def entry(name, bound, work, unit, peak, note=None, key=None):
            ms_ = rows[key or name][0]
            ach = work / (ms_ * 1e-3) / (1e9 if unit == "GB/s" else 1e12)
            e = {"kernel": name, "us": ms_ * 1e3, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                 "frac": (ach / peak) if peak else None, "launches": rows[key or name][1]}
            if note:
                e["note"] = note
            return e

        agg_flops = 2.0 * R * 320 * 1024
        table = [
            entry("hpr_select_kernel (hidden point removal, both problems of the batch)", "alu", B * 12.0 * (n_all + syn.nm + 5 * TRAIN_N),
                  "GB/s", None, "integer/fp64 issue bound, no closed-form roofline (SURVEY 8d): GB/s on the compulsory bytes only",
                  key="hpr_select (both problems)"),
            entry("knn 64-ch layer (tcgen05 Gram screen + exact fp32 re-rank)", "fp32-alu", 2.0 * B * TRAIN_N * TRAIN_N * 64, "TFLOP/s",
                  FP32_FMA_PEAK_TFLOPS, "algorithmic FLOPs of the reference's batched matmul (tf_util.py:613-618)", key="L2 knn (tensor-core part)"),
            entry("knn xyz layer", "fp32-alu", 2.0 * B * TRAIN_N * TRAIN_N * 3, "TFLOP/s", FP32_FMA_PEAK_TFLOPS, key="L1 knn (xyz, tensor-core part)"),
            entry("EdgeConv projection GEMM 64->256 (tf32x3)", "tensor", 2.0 * R * 256 * 64, "TFLOP/s", tf32["tf32_tflops"], key="L4 proj gemm"),
            entry("edge_stats 128 ch", "hbm", R * (256 + k) * 4.0, "GB/s", hbm, key="L4 edge_stats"),
            entry("edge_apply 128 ch", "hbm", R * (256 + k + 2 * 128) * 4.0, "GB/s", hbm, key="L4 edge_apply"),
            entry("edge_bwd_reduce 128 ch", "hbm", R * (256 + k + 128) * 4.0, "GB/s", hbm,
                  key="L4 edge_bwd_reduce (replaced by edge_bwd_stats)" if os.environ.get("CLOUDAAE_EDGE_REC", "0") == "1" else "L4 edge_bwd_reduce"),
            entry("edge_bwd_apply 128 ch", "hbm", R * (256 + k + 128 + 256) * 4.0, "GB/s", hbm, key="L4 edge_bwd_apply"),
            entry("dgcnn_agg forward GEMM 32768x1024x320 (tf32x3, fused BN statistics)", "tensor", agg_flops, "TFLOP/s", tf32["tf32_tflops"],
                  "algorithmic FLOPs; the split-precision product issues 3x that on the tensor pipe (parity: DESIGN 4.2)",
                  key="agg gemm fwd (+stats)"),
            entry("dgcnn_agg data-gradient GEMM (tf32)", "tensor", agg_flops, "TFLOP/s", tf32["tf32_tflops"], key="agg dgrad gemm"),
            entry("dgcnn_agg weight-gradient GEMM (tf32)", "tensor", agg_flops, "TFLOP/s", tf32["tf32_tflops"], key="agg wgrad gemm"),
            entry("bn_act_pool (134 MB pre-activation; records the ReLU-mask statistics of the backward pass)", "hbm", R * 1024 * 4.0, "GB/s", hbm, key="agg bn_act_pool"),
            entry("dgcnn_agg BN backward (coefficients from the pool pass + in-place apply)", "hbm", 2.0 * R * 1024 * 4.0, "GB/s", hbm,
              key="agg bn_bwd (finalize + apply)"),
            entry("nn_distance forward 1024x1024", "fp32-alu", 16.0 * 1024 * 1024 * B, "TFLOP/s", FP32_FMA_PEAK_TFLOPS,
                  "reference two-pass count 16nm FLOP per cloud pair", key="nn_distance fwd"),
            entry("nn_distance backward", "hbm", 32.0 * 2048 * B, "GB/s", hbm, key="nn_distance bwd"),
            entry("FC stack forward (decoder + pose heads)", "hbm", 26.2e6, "GB/s", hbm, "weight bytes, read once", key="fc_fwd"),
            entry("FC stack backward", "hbm", 2 * 26.2e6 + 27.75e6, "GB/s", hbm, "weights read for dgrad, activations for wgrad, gradient written", key="fc_bwd"),
            entry("adam_tf_kernel", "hbm", 7 * 27.75e6, "GB/s", hbm, key="adam"),
        ]
        stage_ms = {n_: rows[n_][0] for n_ in ("synthesis", "prepare_input", "encoder_fwd", "fc_fwd", "losses+chamfer_bwd", "fc_bwd",
                                               "encoder_bwd", "adam", "whole_step")}
        dom = max(table, key=lambda e: e["us"])                                 # by device time, whatever its bound
        closed = max((e for e in table if e["frac"] is not None), key=lambda e: e["us"])
        traffic, traffic_src = ncu_traffic("gemm_tf32_persist_kernel")
        # SURVEY 8(d) composite ceiling of the model part of one 128-segment step: 140 GFLOP / TF32 peak + 5.4 GFLOP /
        # FP32 peak + 0.55 GB / HBM; synthesis has no closed form and is inside the measured step
        ceiling_ms = 140e9 / (tf32["tf32_tflops_sustained"] * 1e12) * 1e3 + 5.4e9 / (FP32_FMA_PEAK_TFLOPS * 1e12) * 1e3 + \
            0.55e9 / (hbm * 1e9) * 1e3
        roofline = {"kernel": closed["kernel"], "bound": closed["bound"], "achieved": closed["achieved"], "peak": closed["peak"],
                    "unit": closed["unit"], "frac": closed["frac"], "traffic": traffic, "traffic_source": traffic_src,
                    "ms_per_launch": closed["us"] / 1e3, "algorithmic_flops_per_launch": agg_flops,
                    # what the tensor pipe executes for it: the split-precision product is three MMAs per k-step
                    # (A B + A_lo B + A B_lo) — reported beside the algorithmic figure, not instead of it
                    "tensor_pipe": ({"mma_flops_per_launch": 3.0 * agg_flops, "achieved": 3.0 * closed["achieved"],
                                     "frac": 3.0 * closed["frac"], "unit": "TFLOP/s"} if "tf32x3" in closed["kernel"] else None),
                    "peak_source": f"measured in this run: {tf32['how']} -> burst {tf32['tf32_tflops']:.0f} / sustained "
                                   f"{tf32['tf32_tflops_sustained']:.0f} TFLOP/s; HBM from MEASURED_PEAKS.json ({peaks['source']})",
                    "dominant_kernel_by_time": {"kernel": dom["kernel"], "us": dom["us"], "bound": dom["bound"],
                                                "share_of_serial_stage_sum": dom["us"] / 1e3 / sum(stage_ms[n_] for n_ in stage_ms if n_ != "whole_step")},
                    "step": {"achieved_segments_per_s": B / (ms * 1e-3), "ceiling_segments_per_s": B / (ceiling_ms * 1e-3),
                             "ceiling_ms": ceiling_ms, "frac": ceiling_ms / ms,
                             "note": "SURVEY 8(d) composite roofline of the model part (synthesis excluded from the ceiling, included in the step)"},
                    "kernels": table}
        synthesis_kernel = table[0]
    return {
        "metric": TRAIN_METRIC, "value": B * world / (ms * 1e-3), "unit": "segments/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (dgcnn_agg contractions: tf32 multiply, f32 accumulate)",
        "data": "synthetic: committed YCB model fixture x fixture pose records, Philox occluders/noise, random-init weights",
        "config": train_config(world, {"cuda_graph": True,
                                       "synthesis": "on-line, inside the timed step" +
                                       ("; two CUDA graphs on two streams with a 2-slot hand-over queue: the synthesis "
                                        "enqueued by step i feeds step i+2 (parallel map + prefetch in the reference)"
                                        if mode not in ("0", "1") else
                                        "; batch i+1 is synthesized next to the train step of batch i (prefetch 1, "
                                        "as tf.data prefetch(1) in the reference)" if pipelined else ""),
                                       "l2": "per-step working set (~0.5 GB of activations) exceeds the 126 MB L2; no flush"}),
        "roofline": roofline, "synthesis_kernel": synthesis_kernel,
        "stage_ms": dict(stage_ms, train_step_pipelined=ms),
        "losses_last_step": losses,
        "repeat_ms_per_step": [round(x, 5) for x in repeats],   # [0] = the contract-timed run; spread of 5 x K steps
        "spread": {"min_ms": min(repeats), "max_ms": max(repeats), "rel": (max(repeats) - min(repeats)) / min(repeats)},
        "e2e": {"value": B * world / (e2e_s / args.steps), "unit": "segments/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16},
        "gpu_launches": int(tr.launches_per_step) * args.steps, "clocks": clocks,
    }


# ----------------------------------------------------------------------------------------------
def _ref_ops_pass(clouds, pred, target, threads):
    """The reference's CPU implementation of the same pass, batch split over `threads` host threads.
    nn_distance fwd+bwd: the reference's own NnDistance/NnDistanceGrad CPU OpKernels (oracle/_ref);
    FPS + gather: oracle port of the CUDA kernel (the reference has no CPU FPS kernel)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import ops as O
    b = clouds.shape[0]
    use_ref = O.have_ref()
    g = np.full((1, OPS_CH), 1.0 / (b * OPS_CH), np.float32)

    def one(i):
        idx = O.fps(clouds[i:i + 1], OPS_M, threads=1)
        O.gather(clouds[i:i + 1], idx)
        if use_ref:
            d1, i1, d2, i2 = O.ref_cpu_nn_distance(pred[i:i + 1], target[i:i + 1])
            O.ref_cpu_nn_distance_grad(pred[i:i + 1], target[i:i + 1], g, i1, g, i2)
        else:
            d1, i1, d2, i2 = O.nn_distance(pred[i:i + 1], target[i:i + 1], "cpu")
            O.nn_distance_grad(pred[i:i + 1], target[i:i + 1], g, i1, g, i2)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(b)))
    return "reference" if use_ref else "port"


def cpu_baseline_ops(sample_clouds: int):
    threads = os.cpu_count() or 1
    b = max(sample_clouds, threads)
    clouds, pred, target = ops_inputs(b, seed=12345)
    _ref_ops_pass(clouds[:threads], pred[:threads], target[:threads], threads)  # warm
    t0 = time.perf_counter()
    reps = 0
    while True:
        kind = _ref_ops_pass(clouds, pred, target, threads)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 20:
            break
    dt = time.perf_counter() - t0
    return {"value": b * reps / dt, "unit": "segments/s", "cores": threads, "kind": kind,
            "sample": f"{reps} x {b} clouds of the same pass (FPS via oracle port; nn_distance fwd+bwd via "
                      f"{'the reference CPU OpKernels' if kind == 'reference' else 'the oracle port'}), "
                      f"batch split over {threads} threads"}


def run_reference_ops(args):
    """--impl reference: the reference's CPU path on this box's host cores, same config/metric."""
    threads = os.cpu_count() or 1
    b = OPS_B
    clouds, pred, target = ops_inputs(b, seed=0)
    for _ in range(max(1, min(args.warmup, 2))):
        _ref_ops_pass(clouds, pred, target, threads)
    t0 = time.perf_counter()
    kind = "port"
    for _ in range(args.steps):
        kind = _ref_ops_pass(clouds, pred, target, threads)
    dt = time.perf_counter() - t0
    value = b * args.steps / dt
    sample = f"{args.steps} steps x {b} clouds, full pass per step, batch split over {threads} host threads"
    return {
        "impl": "reference",
        "metric": "segments/sec (tf_ops microbench pass: FPS 2048->256 + gather + nn_distance fwd+bwd 1024x1024)",
        "value": value, "unit": "segments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (committed YCB model fixture x fixture poses)",
        "config": {"workload": "tf_ops microbench, BASELINE.json configs[1]", "batch_per_gpu": b,
                   "fps": [OPS_N, OPS_M], "nn_distance": [OPS_CH, OPS_CH]},
        "cpu_baseline": {"value": value, "unit": "segments/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ----------------------------------------------------------------------------------------------
def _ref_train_setup(b):
    import torch

    from oracle import model_ref as MR
    params = {k: v.requires_grad_(not k.endswith(("ema_mean", "ema_var")))
              for k, v in MR.init_params(MR.DGCNN_LAYERS, seed=0, perturb=False).items()}
    names = [k for k in params if params[k].requires_grad]
    m = {k: torch.zeros_like(params[k]) for k in names}
    v = {k: torch.zeros_like(params[k]) for k in names}
    return params, names, m, v


def _synth_one(args):
    """Per-sample synthesis exactly as the reference's tf.data map chain (one sample per call)."""
    from oracle import synthesis as S
    model, ax, tr, seed = args
    rng = np.random.default_rng(seed)
    P = S.transform_object_model(model[None], ax[None], tr[None])
    occ = S.spherical_occluder(tr[None, 2], rng.standard_normal((1, 2, 3)), rng.standard_normal((1, 2, 200, 3)))
    fl, org = S.spherical_flip(np.concatenate([P, occ], 1))
    vis, _, _ = S.convex_hull_visible(fl, org)
    fl2, org2 = S.spherical_flip(P)
    vis2, _, _ = S.convex_hull_visible(fl2, org2)
    return vis[0, :TRAIN_N], vis2[0, :4 * TRAIN_N]


def run_reference_train(args):
    """--impl reference: the reference's CPU path for the same step on this box's host cores —
    NumPy + scipy.spatial.ConvexHull synthesis (the reference's own library call) in a process pool,
    the literal TF graph restated in torch-CPU fp32 (all threads), chamfer through the reference's
    own CPU NnDistance/NnDistanceGrad OpKernels when oracle/_ref is built, TF-formula Adam.
    A step is the FULL 128-segment batch (the GPU arm's config) whenever K of them fit in ~4 minutes on this box
    (measured on the warm-up step); otherwise the timed steps take the largest of 64 / 32 / 16 segments that fits and
    the line says so (`config.reference_sample_per_step`, `cpu_baseline.sample`)."""
    import multiprocessing as mp

    import torch

    from oracle import model_ref as MR
    from oracle import ops as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b_full = TRAIN_B
    models = np.load(os.path.join(ROOT, "tests", "golden", "ycb_models_xyz.npy"))
    params, names, m, v = _ref_train_setup(b_full)
    use_ref = O.have_ref()

    class Chamfer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, pred, label):
            fn = O.ref_cpu_nn_distance if use_ref else (lambda x, y: O.nn_distance(x, y, "cpu", threads=cores))
            d1, i1, d2, i2 = fn(pred.detach().numpy(), label.numpy())
            ctx.save = (pred.detach().numpy(), label.numpy(), i1, i2)
            return torch.from_numpy(d1), torch.from_numpy(d2)

        @staticmethod
        def backward(ctx, g1, g2):
            x1, x2, i1, i2 = ctx.save
            fn = O.ref_cpu_nn_distance_grad if use_ref else O.nn_distance_grad
            gx1, _ = fn(x1, x2, g1.numpy(), i1, g2.numpy(), i2)
            return torch.from_numpy(gx1), None

    pool = mp.get_context("fork").Pool(cores)
    batches = pose_batches(b_full, seed=0)

    def step(i, t_idx, b):
        bt = {k_: v_[:b] for k_, v_ in batches[i % len(batches)].items()}
        jobs = [(models[bt["class_id"][k]], bt["axisangle"][k], bt["translation"][k], 1000 * i + k) for k in range(b)]
        res = pool.map(_synth_one, jobs)
        vis = torch.from_numpy(np.stack([r[0] for r in res])); tgt = torch.from_numpy(np.stack([r[1] for r in res]))
        noise = torch.randn(b, TRAIN_N, 3) * (0.004 / 3.0)
        x, mean = MR.prepare_input(vis, torch.from_numpy(bt["class_id"]), noise, num_point=TRAIN_N)
        recon, rot, trans_res, _ = MR.get_model_dgcnn_mean_6d(x, params, True, True, 10, MR.bn_decay_schedule(t_idx, b))
        d1, d2 = Chamfer.apply(recon + mean.unsqueeze(1), tgt)
        chamfer = (d1 + d2).mean()
        tl, _ = MR.get_translation_error(trans_res + mean, torch.from_numpy(bt["translation"]))
        rl, _ = MR.get_rotation_error(rot.double(), torch.from_numpy(bt["axisangle"]).double())
        total = 1000 * chamfer + 10 * tl + rl.float()
        for k in names:
            params[k].grad = None
        total.backward()
        with torch.no_grad():
            for k in names:
                g = params[k].grad if params[k].grad is not None else torch.zeros_like(params[k])
                p_new, m[k], v[k] = MR.adam_step(params[k], g, m[k], v[k], t_idx + 1)
                params[k].copy_(p_new)
        return float(total.item())

    t_idx = 0
    step(0, t_idx, 16); t_idx += 1                       # library warm-up (thread pools, Qhull import in the workers)
    tw = time.perf_counter()
    step(0, t_idx, b_full); t_idx += 1                   # the warm-up step proper: one full 128-segment step, timed
    per_full = time.perf_counter() - tw
    budget = float(os.environ.get("CLOUDAAE_REF_BUDGET_S", "240"))
    b = b_full
    while b > 16 and per_full * (b / b_full) * args.steps > budget:
        b //= 2
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i + 1, t_idx, b); t_idx += 1
    dt = time.perf_counter() - t0
    pool.close()
    value = b * args.steps / dt
    kind = "port"  # TensorFlow 1.12 is not installable here; only the chamfer op is the reference's own binary
    sample = (f"{args.steps} steps x {b} segments per step ({'the full batch' if b == TRAIN_B else 'a bounded sample of the 128-segment batch: ' + str(args.steps) + ' full steps would take ' + format(per_full * args.steps, '.0f') + ' s here'}; "
              f"one full 128-segment step measured at {per_full:.2f} s = {TRAIN_B / per_full:.1f} segments/s): scipy-Qhull synthesis in a "
              f"{cores}-process pool, torch-CPU fp32 restatement of the TF graph on {cores} threads, chamfer via "
              f"{'the reference CPU OpKernels (oracle/_ref)' if use_ref else 'the oracle port'}, Adam")
    return {
        "impl": "reference", "metric": TRAIN_METRIC, "value": value, "unit": "segments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (same fixtures as the GPU arm)",
        "config": train_config(1, {"reference_sample_per_step": b, "reference_full_step_s": per_full}),
        "cpu_baseline": {"value": value, "unit": "segments/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference_infer(args):
    """--impl reference --workload infer: the evaluation graph (evaluate_cloudAAE_ycbv.py:437-474) on this box's
    host cores — torch-CPU fp32 restatement of the TF graph with moving-average batch norm, FPS 1024 -> 256 by the
    oracle's C restatement of the reference CUDA kernel (the reference has no CPU FPS op), chamfer through the
    reference's own CPU OpKernel when oracle/_ref is built.  Segments are synthesized before the timed region.
    Each step is a bounded SAMPLE (64 segments) of the 4096-segment list."""
    import multiprocessing as mp

    import torch

    from oracle import model_ref as MR
    from oracle import ops as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = 64
    models = np.load(os.path.join(ROOT, "tests", "golden", "ycb_models_xyz.npy"))
    params, _, _, _ = _ref_train_setup(b)
    for k in params:
        if k.endswith("ema_var"):
            params[k].fill_(1.0)
    use_ref = O.have_ref()
    bt = pose_batches(b, seed=0, pool=1)[0]
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_synth_one, [(models[bt["class_id"][k]], bt["axisangle"][k], bt["translation"][k], k) for k in range(b)])
    vis = torch.from_numpy(np.stack([r[0] for r in res])); tgt = np.stack([r[1][:TRAIN_N] for r in res])
    cls = torch.from_numpy(bt["class_id"])

    def step():
        with torch.no_grad():
            x, mean = MR.prepare_input(vis, cls, torch.zeros(b, TRAIN_N, 3), num_point=TRAIN_N)
            recon, rot, trans_res, _ = MR.get_model_dgcnn_mean_6d(x, params, False, False, 10, None)
            recon = (recon + mean.unsqueeze(1)).numpy()
            idx = O.fps(recon, TRAIN_N, threads=cores)
            sub = O.gather(recon, idx)
            fn = O.ref_cpu_nn_distance if use_ref else (lambda a, c: O.nn_distance(a, c, "cpu", threads=cores))
            d1, _, d2, _ = fn(sub, tgt)
            MR.get_translation_error(trans_res + mean, torch.from_numpy(bt["translation"]))
            MR.get_rotation_error(rot.double(), torch.from_numpy(bt["axisangle"]).double())
        return float((d1 + d2).mean())

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = b * args.steps / dt
    sample = (f"{args.steps} steps x {b}-segment sample of the {INFER_TOTAL}-segment list: torch-CPU fp32 restatement of the "
              f"TF graph (moving-average BN) on {cores} threads, FPS by the oracle's C restatement, chamfer via "
              f"{'the reference CPU OpKernel (oracle/_ref)' if use_ref else 'the oracle port'}")
    return {
        "impl": "reference", "metric": "inference segments/sec", "value": value, "unit": "segments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (same fixtures as the GPU arm)",
        "config": {"workload": "batched inference over all 21 YCB classes, BASELINE.json configs[4]",
                   "segments": INFER_TOTAL, "reference_sample_per_step": b, "num_point": TRAIN_N},
        "cpu_baseline": {"value": value, "unit": "segments/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def cpu_baseline_train(seconds: float = 12.0):
    """cpu_baseline leg of the default run: the same reference step, run for ~`seconds` s on rank 0."""
    class A:  # minimal args
        gpus, warmup = 1, 1
        steps = 2
    t0 = time.perf_counter()
    res = run_reference_train(A)
    res["cpu_baseline"]["wall_s"] = time.perf_counter() - t0
    return res["cpu_baseline"]


# ----------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------
# Workload "infer": BASELINE.json configs[4] — batched inference over all 21 YCB classes, 4096 segments,
# sharded over the ranks with no collective (evaluate_cloudAAE_ycbv.py:405-477 per batch: mean-normalise,
# get_model_dgcnn_mean_6d with moving-average BN, FPS 1024 -> 256 of the reconstruction, chamfer, pose errors).
INFER_TOTAL, INFER_B = 4096, 128


def run_ours_infer(args, rank, world, local_rank, stages=True):
    import torch
    import torch.distributed as dist

    from cloudaae_b200 import _capi
    from cloudaae_b200 import evaluate_cloudAAE_ycbv as EV
    from cloudaae_b200.inference import CloudAAEInference
    from cloudaae_b200.models import pointnet_ycb_23_decoder_4 as M
    from cloudaae_b200.parallel import shard_range
    from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, N = INFER_B, TRAIN_N
    lo, hi = shard_range(INFER_TOTAL, rank, world)
    nb = (hi - lo + B - 1) // B
    z = np.load(os.path.join(ROOT, "tests", "golden", "ycb_poses.npz"))
    per = len(z["class_id"]) // 21
    ids = np.arange(lo, lo + nb * B) % INFER_TOTAL
    cls_np = (ids % 21).astype(np.int32)                      # classes cycling 0..20
    rec = cls_np.astype(np.int64) * per + (ids // 21) % per
    assert (z["class_id"][rec] == cls_np).all()
    models = load_models_xyz(device=dev)
    syn = SegmentSynthesizer(models, B, N, seed=99 + rank)
    v = M.Variables(M.dgcnn_layers(N, 3 + M.NUM_CLASS), device=dev, seed=0)
    v.ema.fill_(0.0)
    for name, (o, shape) in v.ema_index.items():              # untrained weights: unit moving variance
        if name.endswith("ema_var"):
            v.ema[o:o + int(np.prod(shape))] = 1.0
    inf = CloudAAEInference(v, batch_size=B, num_point=N)
    seg = torch.empty(nb, B, N, 3, device=dev); tgt = torch.empty(nb, B, N, 3, device=dev)
    cls = torch.from_numpy(cls_np).to(dev).view(nb, B)
    tl = torch.from_numpy(z["translation"][rec].astype(np.float32)).to(dev).view(nb, B, 3)
    ax = torch.from_numpy(z["axisangle"][rec].astype(np.float32)).to(dev).view(nb, B, 3)
    for i in range(nb):                                        # synthetic segments, resident in HBM
        vis, target, noise = syn.synthesize(cls[i], ax[i], tl[i])
        seg[i] = vis + noise
        tgt[i] = target[:, :N]
    torch.cuda.synchronize()

    graph = os.environ.get("CLOUDAAE_INFER_GRAPH", "1") != "0"
    st = inf.capture() if graph else None

    def one_pass():
        for i in range(nb):
            if graph:   # one graph launch per batch; the inputs are device-to-device copies into its static buffers
                st["segment"].copy_(seg[i]); st["class_id"].copy_(cls[i]); st["target"].copy_(tgt[i])
                st["translation"].copy_(tl[i]); st["axisangle"].copy_(ax[i])
                inf.replay()
            else:
                inf.forward(seg[i], cls[i], tgt[i], tl[i], ax[i])

    for _ in range(args.warmup):
        one_pass()
    torch.cuda.synchronize()
    c0 = _capi.COUNTER[0]
    one_pass()
    launches_per_pass = (_capi.COUNTER[0] - c0) if not graph else nb * inf.launches_per_batch
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        one_pass()
    e.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: pinned host segments -> H2D -> forward -> D2H of the pose outputs, every batch
    seg_h, cls_h = seg.cpu().pin_memory(), cls.cpu().pin_memory()
    seg_d, cls_d = torch.empty(B, N, 3, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
    rot_h, tr_h = torch.empty(B, 3).pin_memory(), torch.empty(B, 3).pin_memory()

    def e2e_pass():
        for i in range(nb):
            if graph:
                st["segment"].copy_(seg_h[i], non_blocking=True); st["class_id"].copy_(cls_h[i], non_blocking=True)
                out = inf.replay()
            else:
                seg_d.copy_(seg_h[i], non_blocking=True); cls_d.copy_(cls_h[i], non_blocking=True)
                out = inf.forward(seg_d, cls_d)
            rot_h.copy_(out["rot_pred"], non_blocking=True); tr_h.copy_(out["trans_pred"], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    e2e_pass()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_pass()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank != 0:
        return None

    ms = total_ms / args.steps
    n_seg = nb * B * world

    class RefArgs:
        gpus, warmup, steps = 1, 1, 3
    cpu_ref = run_reference_infer(RefArgs)["cpu_baseline"]
    base = {
        "cpu_baseline": cpu_ref,
        "metric": "inference segments/sec", "value": n_seg / (ms * 1e-3), "unit": "segments/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 (tensor-core contractions: split-precision 3xTF32, f32 accumulate)",
        "data": "synthetic: committed YCB model fixture x fixture pose records through the on-line synthesis; random-init weights",
        "config": {"workload": "batched inference over all 21 YCB classes, BASELINE.json configs[4]",
                   "segments": n_seg, "batch_per_forward": B, "num_point": N, "cuda_graph": graph,
                   "parallelism": f"segment list sharded over {world} rank(s), no collective",
                   "l2": "each forward streams ~0.3 GB of activations (> 126 MB L2); no flush"},
        "e2e": {"value": n_seg / (e2e_s / args.steps), "unit": "segments/s",
                "h2d_bytes_per_step": nb * (B * N * 12 + B * 4), "d2h_bytes_per_step": nb * B * 24},
        "gpu_launches": int(launches_per_pass) * args.steps, "clocks": clocks,
    }
    if not stages:
        return base

    # ---- the stages either side of the network, timed on their own (rank 0)
    stream = torch.cuda.current_stream()

    def timeit(fn, iters=5):
        fn(); torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(iters):
            fn()
        b_.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / iters

    import random
    from cloudaae_b200.data import synthetic_frames as OE
    posed = syn.points[:12, :syn.nm].cpu().numpy()            # the last synthesized batch, camera frame
    last_cls = cls[nb - 1].cpu().numpy()
    frames_d, frames_l, fos, cos = [], [], [], []
    for f in range(4):
        cl = [int(c) for c in last_cls[f * 3:f * 3 + 3]]
        d_, l_ = OE.render_frame(posed[f * 3:f * 3 + 3], cl, splat=2, seed=f)
        frames_d.append(d_); frames_l.append(l_); fos += [f] * 3; cos += cl
    fe = EV.SegmentFrontEnd(torch.from_numpy(np.stack(frames_d)).to(dev), torch.from_numpy(np.stack(frames_l)).to(dev),
                            torch.from_numpy(np.tile(OE.YCBV_INTRINSICS, (4, 1))).to(dev),
                            torch.full((21,), 0.2, device=dev), cap=49152)
    t_front = timeit(lambda: fe.run(fos, cos, N, rng=random.Random(0)))
    fr = fe.run(fos, cos, N, rng=random.Random(0))
    src6 = torch.cat([models, torch.zeros_like(models)], dim=2).contiguous()
    T0 = EV.pose_to_matrix(ax[0], tl[0])
    T0[:, :3, 3] += 0.003
    t_icp = timeit(lambda: EV.icp_refine(src6, seg[0], T0, source_of_seg=cls[0]))
    _, fit, rmse, iters = EV.icp_refine(src6, seg[0], T0, source_of_seg=cls[0])
    return dict(base, **{
        "front_end": {"what": "12 (frame, class) segments from 4 synthetic 480x640 frames: extract + mean filter, radius "
                              "outliers, two float64 FPS_random to 256 points (SURVEY 8f rank 2)",
                      "ms": t_front, "points_after_filter": fr["num_point_after_filter"].tolist()},
        "icp": {"what": "128 segments x 10 registration_icp rounds, model 2048 pts -> segment 256 pts, one launch "
                        "(SURVEY 8f rank 4)", "ms": t_icp, "mean_iterations": float(iters.float().mean()),
                "mean_fitness": float(fit.mean()), "mean_inlier_rmse": float(rmse.mean())},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            fn = {"auto": run_reference_train, "train": run_reference_train,
                  "infer": run_reference_infer}.get(args.workload, run_reference_ops)
            print(json.dumps(fn(args)), flush=True)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        from cloudaae_b200.parallel import init_nccl
        init_nccl(local_rank)   # NCCL communicator capped at 8 CTAs: the step is bound by SM work
    if args.workload in ("auto", "train"):
        result = run_ours_train(args, rank, world, local_rank)
        if rank == 0 and world == 1 and os.environ.get("CLOUDAAE_BENCH_LIGHT", "0") != "1":
            # the metric also asks for FPS / nn_distance achieved GB/s: run the tf_ops microbench
            # (BASELINE configs[1]) on rank 0 after the timed region and attach its kernel table
            class OpsArgs:
                steps, warmup = 30, 3
            ops = run_ours_ops(OpsArgs, 0, 1, local_rank, with_cpu_baseline=True)
            result["ops_microbench"] = {k: ops[k] for k in ("metric", "value", "unit", "ms_per_step", "kernels", "roofline", "e2e",
                                                            "cpu_baseline", "reference_kernels_sm100a", "config")}
            # BASELINE configs[4] (batched inference, 4096 segments) as a compact block of the same line
            class InfArgs:
                steps, warmup = 3, 3
            inf = run_ours_infer(InfArgs, 0, 1, local_rank, stages=False)
            result["infer"] = {k: inf[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "cpu_baseline", "config",
                                                   "gpu_launches")}
            result["cpu_baseline"] = cpu_baseline_train()
    elif args.workload == "infer":
        result = run_ours_infer(args, rank, world, local_rank)
    else:
        result = run_ours_ops(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        # NCCL communicators referenced by a live CUDA graph make destroy_process_group() block; the
        # timed region is over and the line is printed, so leave without tearing the group down.
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
