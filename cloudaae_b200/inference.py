"""Batched inference — the GPU graph of evaluate_cloudAAE_ycbv.py (:405-477) for many segments.

Per batch: mean-normalise the input segment (:437-439), get_model_dgcnn_mean_6d with both BN flags
False (:442-444, moving averages), add the mean back (:446-447), FPS 4N -> N on the reconstruction
fused with gather_point (:450), chamfer against the first N target points (:452), translation and
rotation errors (:456-474).  Segments are independent, so a segment list is sharded by rank with no
collective (BASELINE config 5: 4096 segments over 8 GPUs).
"""
from __future__ import annotations

import torch

from . import _capi
from .models.pointnet_ycb_23_decoder_4 import NUM_CLASS, Variables, _Engine
from .parallel import shard_range


class CloudAAEInference:
    def __init__(self, variables: Variables, batch_size: int, num_point: int = 256, k_neighbor: int = 10,
                 precision: str | None = None):
        self.v = variables
        self.B, self.N = batch_size, num_point
        self.M = 4 * num_point
        self.dev = variables.device
        self.lib = _capi.lib()
        self.engine = _Engine(variables, "dgcnn", batch_size, num_point, 3 + NUM_CLASS, k_neighbor, precision=precision)
        f32 = dict(dtype=torch.float32, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        B, N, M = self.B, self.N, self.M
        self.x = torch.empty(B, N, 3 + NUM_CLASS, **f32)
        self.mean = torch.empty(B, 3, **f32)
        self.recon = torch.empty(B, M, 3, **f32)
        self.fps_idx = torch.empty(B, N, **i32)
        self.recon_fps = torch.empty(B, N, 3, **f32)
        self.dist1 = torch.empty(B, N, **f32); self.dist2 = torch.empty(B, N, **f32)
        self.idx1 = torch.empty(B, N, **i32); self.idx2 = torch.empty(B, N, **i32)
        self.per_rot = torch.empty(B, dtype=torch.float64, device=self.dev)
        self.per_trans = torch.empty(B, **f32)
        self.scratch = torch.empty(B, 3, **f32); self.scratch2 = torch.empty(B, 3, **f32)
        self.trans_pred = torch.empty(B, 3, **f32)

    def _c(self, name, *args):
        _capi.check(getattr(self.lib, name)(*args, torch.cuda.current_stream(self.dev).cuda_stream), name)

    def forward(self, segment, class_id, target=None, translation=None, axisangle=None):
        """segment f32[B,>=N,3] (first N rows used), class_id i32[B]; optional labels: target f32[B,N,3]
        (chamfer), translation / axisangle f32[B,3] (pose errors).  Returns a dict of device tensors
        (views of reused buffers): recon [B,4N,3], recon_fps [B,N,3], fps_idx, rot_pred, trans_pred and,
        when labels are given, chamfer [B,N], trans_err [B], rot_err [B] (float64)."""
        B, N, M = self.B, self.N, self.M
        p = _Engine._p
        assert segment.is_contiguous() and segment.shape[0] == B and segment.shape[1] >= N
        self._c("caae_prepare_input", B, N, segment.shape[1], p(segment), None, p(class_id), NUM_CLASS, p(self.x),
                p(self.mean))
        recon, rot, trans, emb, _ = self.engine.forward(self.x, False, False, None)
        self._c("caae_add_cloud_vec", B, M, p(recon), p(self.mean), p(self.recon))
        self._c("caae_fps_gather", B, M, N, p(self.recon), None, p(self.fps_idx), p(self.recon_fps))
        out = {"recon": self.recon, "recon_fps": self.recon_fps, "fps_idx": self.fps_idx, "rot_pred": rot,
               "embedding": emb}
        if target is not None:
            assert target.is_contiguous() and target.shape == (B, N, 3)
            self._c("caae_nn_distance", B, N, p(self.recon_fps), N, p(target), p(self.dist1), p(self.idx1),
                    p(self.dist2), p(self.idx2))
            out["chamfer"] = self.dist1 + self.dist2
        tl = translation if translation is not None else self.scratch
        al = axisangle if axisangle is not None else self.scratch
        self._c("caae_pose_losses", B, p(rot), p(al), p(trans), p(self.mean), p(tl), 1.0, 1.0, p(self.per_rot),
                p(self.per_trans), p(self.scratch2), p(self.scratch2), p(self.trans_pred))
        out["trans_pred"] = self.trans_pred
        if translation is not None:
            out["trans_err"] = self.per_trans
        if axisangle is not None:
            out["rot_err"] = self.per_rot
        return out


    # ------------------------------------------------------------------ CUDA graph
    def capture(self, with_target: bool = True, with_pose_labels: bool = True, seg_points: int | None = None):
        """Capture `forward` on static input buffers (the ~50 launches of a batch become one graph launch:
        eval-mode batches are launch-bound otherwise).  Returns the static inputs dict {'segment', 'class_id',
        'target', 'translation', 'axisangle'} (absent labels are None); fill them, then `replay()` returns the
        same output dict as `forward` (views of reused buffers)."""
        B, N = self.B, self.N
        f32 = dict(dtype=torch.float32, device=self.dev)
        st = {"segment": torch.zeros(B, seg_points or N, 3, **f32),
              "class_id": torch.zeros(B, dtype=torch.int32, device=self.dev),
              "target": torch.zeros(B, N, 3, **f32) if with_target else None,
              "translation": torch.zeros(B, 3, **f32) if with_pose_labels else None,
              "axisangle": torch.zeros(B, 3, **f32) if with_pose_labels else None}
        args = (st["segment"], st["class_id"], st["target"], st["translation"], st["axisangle"])
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.forward(*args)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        before = _capi.COUNTER[0]
        with torch.cuda.graph(self._graph, stream=side):
            self._graph_out = self.forward(*args)
        self.launches_per_batch = _capi.COUNTER[0] - before
        self.static = st
        return st

    def replay(self):
        self._graph.replay()
        return self._graph_out


def run_sharded(infer: CloudAAEInference, segments, class_ids, targets, translations, axisangles, rank: int = 0,
                world: int = 1):
    """Process this rank's contiguous shard of a segment list (device tensors, leading dim = total).
    Returns (start, stop, per-segment chamfer mean, trans_err, rot_err) for the shard.  The tail batch
    is padded by repeating the last segment; no collective is involved."""
    total = segments.shape[0]
    a, b = shard_range(total, rank, world)
    B = infer.B
    cham, terr, rerr = [], [], []
    for s in range(a, b, B):
        e = min(s + B, b)
        sel = torch.arange(s, s + B, device=segments.device).clamp_(max=e - 1)
        out = infer.forward(segments[sel].contiguous(), class_ids[sel].contiguous(), targets[sel].contiguous(),
                            translations[sel].contiguous(), axisangles[sel].contiguous())
        n = e - s
        cham.append(out["chamfer"][:n].mean(dim=1).clone()); terr.append(out["trans_err"][:n].clone())
        rerr.append(out["rot_err"][:n].clone())
    cat = lambda xs, dt: torch.cat(xs) if xs else torch.empty(0, dtype=dt, device=segments.device)  # noqa: E731
    return a, b, cat(cham, torch.float32), cat(terr, torch.float32), cat(rerr, torch.float64)
