"""Evaluation-side stages around the network — drop-ins for the CPU / py_func stages of the reference's
``evaluate_cloudAAE_ycbv.py``, on CUDA tensors, batched over segments.

    get_pointcloud(depth, fx, fy, cx, cy, depth_scaling_factor)            (:164-178)
    segment_mean_distance_filter / segment_not_empty / outlier_removal     (:219-223, :262-281)
    get_outlier_idx(xyz, nb_points, radius, std_ratio)                     (:250-258, open3d radius outliers)
    FPS_random(pts, K, ...)                                                (:230-247, float64, random first index)
    icp_refine(...)                                                        (:606-624, open3d registration_icp loop)

A *segment* is one (frame, class) pair.  ``SegmentFrontEnd`` runs the whole chain for a list of segments
with four launches; the single-segment functions keep the reference's names and argument order.  The
random first index of ``FPS_random`` (``random.randint`` in the reference) is drawn on the host from the
caller's ``random.Random`` so a seeded run is reproducible.  No CPU fallback.
"""
from __future__ import annotations

import random as _random

import torch

from . import _capi
from ._capi import InvalidArgumentError

__all__ = ["get_pointcloud", "get_outlier_idx", "FPS_random", "SegmentFrontEnd", "icp_refine", "pose_to_matrix"]


def _stream(t):
    return _capi.stream_of(t)


def _as_i32(x, device):
    return torch.as_tensor(x, dtype=torch.int32, device=device).contiguous()


class SegmentFrontEnd:
    """depth / label frames -> network-ready segments, all on the GPU.

    frames: depth u16-valued [F,h,w] (torch.uint16 or int16 storage), label uint8 [F,h,w] (one-based class
    labels), intrinsics f32[F,5] = (fx, fy, cx, cy, factor_depth).  ``cap`` bounds the points kept per
    segment (rows past it are dropped; the counts returned are the true ones)."""

    def __init__(self, depth: torch.Tensor, label: torch.Tensor, intrinsics: torch.Tensor,
                 threshold_distance_per_class: torch.Tensor, cap: int = 32768):
        if depth.dim() != 3 or label.shape != depth.shape:
            raise InvalidArgumentError("SegmentFrontEnd expects depth and label of shape (frames, height, width)")
        if depth.dtype not in (torch.uint16, torch.int16) or label.dtype != torch.uint8:
            raise InvalidArgumentError("SegmentFrontEnd expects 16-bit depth and uint8 label")
        if intrinsics.shape != (depth.shape[0], 5) or intrinsics.dtype != torch.float32:
            raise InvalidArgumentError("SegmentFrontEnd expects float32 intrinsics of shape (frames, 5)")
        if threshold_distance_per_class.dim() != 1:
            raise InvalidArgumentError("SegmentFrontEnd expects one distance threshold per class")
        for t, name in ((depth, "depth"), (label, "label"), (intrinsics, "intrinsics"),
                        (threshold_distance_per_class, "threshold_distance_per_class")):
            _capi.require_cuda(t, f"SegmentFrontEnd({name})")
        self.depth, self.label = depth.contiguous(), label.contiguous()
        self.intr = intrinsics.contiguous()
        self.thr = threshold_distance_per_class.to(torch.float32).contiguous()
        self.cap = int(cap)
        self.dev = depth.device
        self.lib = _capi.lib()

    def extract(self, frame_of_seg, class_of_seg, want_org: bool = True):
        """segment_not_empty + the masks of outlier_removal.  Returns a dict: xyz_org [S,cap,3], n_org [S],
        xyz_org_distance_filtered [S,cap,3], pix [S,cap] (pixel ids, for rgb), num_point_after_filter [S],
        mean [S,3]."""
        # host-side lists are range-checked before they reach the kernel (device tensors are the caller's contract)
        for vals, hi, what in ((frame_of_seg, self.depth.shape[0], "frame"), (class_of_seg, self.thr.shape[0], "class")):
            if not isinstance(vals, torch.Tensor) and len(vals) and not all(0 <= int(x) < hi for x in vals):
                raise InvalidArgumentError(f"SegmentFrontEnd.extract: {what} index outside [0, {hi})")
        f = _as_i32(frame_of_seg, self.dev)
        c = _as_i32(class_of_seg, self.dev)
        if f.dim() != 1 or f.shape != c.shape:
            raise InvalidArgumentError("frame_of_seg and class_of_seg must be 1-D and of equal length")
        S, cap = f.shape[0], self.cap
        F, h, w = self.depth.shape
        f32 = dict(dtype=torch.float32, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        out = {
            "xyz_org": torch.zeros(S, cap, 3, **f32) if want_org else None,
            "n_org": torch.zeros(S, **i32) if want_org else None,
            "xyz_org_distance_filtered": torch.zeros(S, cap, 3, **f32),
            "pix": torch.zeros(S, cap, **i32),
            "num_point_after_filter": torch.zeros(S, **i32),
            "mean": torch.zeros(S, 3, **f32),
        }
        p = _capi.ptr
        with torch.cuda.device(self.dev):
            _capi.check(self.lib.caae_segment_extract(
                S, F, h, w, p(f), p(c), p(self.depth), p(self.label), p(self.intr), p(self.thr), cap,
                p(out["xyz_org"]), p(out["n_org"]), p(out["xyz_org_distance_filtered"]), p(out["pix"]),
                p(out["num_point_after_filter"]), p(out["mean"]), _stream(f)), "caae_segment_extract")
        return out

    def radius_outliers(self, xyz: torch.Tensor, n_pts: torch.Tensor, nb_points: int = 100, radius: float = 0.02,
                        min_keep: int = 512):
        """get_outlier_idx for every segment: (inlier_idx i32[S,cap] ascending, n_inlier i32[S])."""
        return _radius_outliers(xyz, n_pts, nb_points, radius, min_keep)

    def run(self, frame_of_seg, class_of_seg, numpoints: int, rng: _random.Random | None = None,
            nb_points: int = 100, radius: float = 0.02):
        """create_tfrecord_dataset's per-segment chain (evaluate…:314-322) up to the network input:
        extract -> radius outliers -> FPS_random on both the distance-filtered and the inlier cloud.
        Returns the reference's element keys ('xyz', 'xyz_inlier', 'xyz_inlier_full', 'xyz_org_distance_filtered',
        'xyz_org', 'num_valid_points_in_segment', 'num_point_after_filter') plus the counts needed to read the
        padded arrays, and 'keep' = the two dataset filters (num_point_after_filter > 100, num_valid >= numpoints)."""
        rng = rng or _random
        e = self.extract(frame_of_seg, class_of_seg)
        flt, n_flt = e["xyz_org_distance_filtered"], e["num_point_after_filter"]
        idx, n_in = _radius_outliers(flt, n_flt, nb_points, radius, 512)
        inl = _gather_rows(flt, idx)
        n_flt_h, n_in_h = n_flt.tolist(), n_in.tolist()
        # the counts are the true ones, the padded arrays hold at most `cap` points in raster order: a larger segment would
        # silently lose its bottom rows and diverge from the reference — refuse instead (size `cap` from the label counts)
        n_max = max(max(n_flt_h, default=0), int(e["n_org"].max()) if e["n_org"] is not None and len(n_flt_h) else 0)
        if n_max > self.cap:
            raise InvalidArgumentError(f"SegmentFrontEnd.run: a segment has {n_max} points but cap = {self.cap}; "
                                       f"construct the front end with cap >= {n_max}")
        # the reference draws random.randint(0, n-1) per FPS_random call, inlier cloud first (evaluate…:288-289)
        first_in = [rng.randint(0, max(n - 1, 0)) for n in n_in_h]
        first_org = [rng.randint(0, max(min(n, self.cap) - 1, 0)) for n in n_flt_h]
        fi, xyz_inlier = _fps_seeded(inl, n_in, _as_i32(first_in, self.dev), numpoints)
        fo, xyz = _fps_seeded(flt, n_flt, _as_i32(first_org, self.dev), numpoints)
        # tf.count_nonzero(inlier_idx): non-zero index VALUES (evaluate…:279) — a kept point 0 is not counted
        has0 = (idx[:, 0] == 0) & (n_in > 0)
        num_valid = n_in - has0.to(torch.int32)
        e.update({"inlier_idx": idx, "n_inlier": n_in, "xyz_inlier_full": inl, "xyz": xyz, "xyz_inlier": xyz_inlier,
                  "FPS_org_idx": fo, "FPS_inlier_idx": fi, "num_valid_points_in_segment": num_valid,
                  "keep": (n_flt > 100) & (num_valid >= numpoints)})
        return e


def _radius_outliers(xyz, n_pts, nb_points, radius, min_keep):
    if xyz.dim() != 3 or xyz.shape[2] != 3 or xyz.dtype != torch.float32:
        raise InvalidArgumentError("get_outlier_idx expects float32 xyz of shape (segments, cap, 3)")
    if nb_points < 0 or not radius >= 0:
        raise InvalidArgumentError("get_outlier_idx expects nb_points >= 0 and radius >= 0")
    _capi.require_cuda(xyz, "get_outlier_idx")
    xyz = xyz.contiguous()
    S, cap, _ = xyz.shape
    n_pts = _as_i32(n_pts, xyz.device)
    flag = torch.empty(S, cap, dtype=torch.uint8, device=xyz.device)
    idx = torch.empty(S, cap, dtype=torch.int32, device=xyz.device)
    n_in = torch.empty(S, dtype=torch.int32, device=xyz.device)
    p = _capi.ptr
    with torch.cuda.device(xyz.device):
        _capi.check(_capi.lib().caae_radius_outlier(S, cap, p(xyz), p(n_pts), int(nb_points), float(radius),
                                                    int(min_keep), p(flag), p(idx), p(n_in), _stream(xyz)),
                    "caae_radius_outlier")
    return idx, n_in


def _gather_rows(xyz, idx):
    S, cap, _ = xyz.shape
    out = torch.empty_like(xyz)
    p = _capi.ptr
    with torch.cuda.device(xyz.device):
        _capi.check(_capi.lib().caae_gather(S, cap, cap, p(xyz), p(idx), p(out), _stream(xyz)), "caae_gather")
    return out


def _fps_seeded(xyz, n_pts, first_idx, k):
    xyz = xyz.contiguous()
    S, cap, _ = xyz.shape
    temp = torch.empty(S, cap, dtype=torch.float64, device=xyz.device)
    out_idx = torch.empty(S, k, dtype=torch.int32, device=xyz.device)
    out_xyz = torch.empty(S, k, 3, dtype=torch.float32, device=xyz.device)
    p = _capi.ptr
    with torch.cuda.device(xyz.device):
        _capi.check(_capi.lib().caae_fps_seeded_f64(S, cap, int(k), p(xyz), p(n_pts), p(first_idx), p(temp),
                                                    p(out_idx), p(out_xyz), _stream(xyz)), "caae_fps_seeded_f64")
    return out_idx, out_xyz


# ---- single-segment functions under the reference's names ---------------------------------------------

def get_pointcloud(depth: torch.Tensor, fx, fy, cx, cy, depth_scaling_factor) -> torch.Tensor:
    """depth u16 [h,w] -> f32 [h*w,3] for ALL pixels (evaluate…:164-178).  Runs the segment kernel with an
    all-ones label (its mask drops zero-depth pixels) and scatters the rows back by pixel id; a zero-depth pixel
    keeps the row (0, 0, 0), which is what the formula yields for depth 0."""
    if depth.dim() != 2:
        raise InvalidArgumentError("get_pointcloud expects a (height, width) depth image")
    _capi.require_cuda(depth, "get_pointcloud")
    h, w = depth.shape
    dev = depth.device
    label = torch.ones(1, h, w, dtype=torch.uint8, device=dev)
    intr = torch.tensor([[fx, fy, cx, cy, depth_scaling_factor]], dtype=torch.float32, device=dev)
    fe = SegmentFrontEnd(depth.reshape(1, h, w), label, intr, torch.full((1,), float("inf"), device=dev), cap=h * w)
    e = fe.extract([0], [0], want_org=False)
    n = int(e["num_point_after_filter"][0])
    out = torch.zeros(h * w, 3, dtype=torch.float32, device=dev)
    out[e["pix"][0, :n].long()] = e["xyz_org_distance_filtered"][0, :n]
    return out


def get_outlier_idx(xyz: torch.Tensor, nb_points: int, radius: float, std_ratio: float = 0.5) -> torch.Tensor:
    """One segment (n,3) or (1,n,3) -> int64 inlier ids (evaluate…:250-258).  std_ratio is unused there too."""
    pts = xyz.reshape(1, -1, 3).to(torch.float32)
    idx, n_in = _radius_outliers(pts, [pts.shape[1]], nb_points, radius, 512)
    return idx[0, : int(n_in[0])].long()


def FPS_random(pts: torch.Tensor, K: int, seq_id=None, frame_id=None, class_id=None, first_idx: int | None = None,
               rng: _random.Random | None = None) -> torch.Tensor:
    """One segment (n,>=3) -> int64 [K] (evaluate…:230-247).  seq_id/frame_id/class_id only feed the
    reference's log lines; `first_idx` (or `rng`) replaces the module-level random.randint."""
    if pts.dim() != 2 or pts.shape[1] < 3:
        raise InvalidArgumentError("FPS_random expects pts of shape (points, >=3)")
    n = pts.shape[0]
    if n == 0:
        raise ValueError("FPS_random: empty segment (random.randint(0, -1) raises in the reference)")
    if K <= 0:
        raise InvalidArgumentError("FPS_random expects a positive K")
    _capi.require_cuda(pts, "FPS_random")
    if first_idx is None:
        first_idx = (rng or _random).randint(0, n - 1)
    xyz = pts[:, 0:3].to(torch.float32).reshape(1, n, 3).contiguous()
    idx, _ = _fps_seeded(xyz, _as_i32([n], pts.device), _as_i32([first_idx], pts.device), K)
    return idx[0].long()


# ---- ICP refinement -----------------------------------------------------------------------------------

def pose_to_matrix(rot_axag: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
    """[rotmat | trans] as float64 4x4 (evaluate…:571-575 + :609-611): Rodrigues of the predicted axis-angle."""
    a = rot_axag.to(torch.float64)
    ang = a.norm(dim=1, keepdim=True)
    ax = a / ang
    K = torch.zeros(a.shape[0], 3, 3, dtype=torch.float64, device=a.device)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -ax[:, 2], ax[:, 1], ax[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 0], -ax[:, 1], ax[:, 0]
    s, c = torch.sin(ang)[:, :, None], torch.cos(ang)[:, :, None]
    R = torch.eye(3, dtype=torch.float64, device=a.device)[None] + s * K + (1 - c) * (K @ K)
    T = torch.zeros(a.shape[0], 4, 4, dtype=torch.float64, device=a.device)
    T[:, :3, :3] = R
    T[:, :3, 3] = trans.to(torch.float64)
    T[:, 3, 3] = 1.0
    return T


def icp_refine(source: torch.Tensor, target: torch.Tensor, init: torch.Tensor, source_of_seg=None,
               radius: float = 0.01, radius_decay: float = 0.9, outer: int = 10, max_iteration: int = 30,
               relative_fitness: float = 1e-6, relative_rmse: float = 1e-6):
    """The ICP loop of evaluate…:615-624 for a batch: `outer` registration_icp calls (point-to-point) with
    the radius shrinking by `radius_decay`.  source f32[nsrc,ns,>=3] (xyz = first three columns, e.g. the
    object models [21,2048,6]; `source_of_seg` i32[B] picks the model of each segment, default = segment
    index), target f32[B,nt,3], init f64[B,4,4].  Returns (T f64[B,4,4], fitness f64[B], inlier_rmse f64[B],
    iterations i32[B])."""
    if source.dim() != 3 or source.shape[2] < 3 or source.dtype != torch.float32:
        raise InvalidArgumentError("icp_refine expects float32 source of shape (models, points, >=3)")
    if target.dim() != 3 or target.shape[2] != 3 or target.dtype != torch.float32:
        raise InvalidArgumentError("icp_refine expects float32 target of shape (batch, points, 3)")
    B = target.shape[0]
    if init.shape != (B, 4, 4) or init.dtype != torch.float64:
        raise InvalidArgumentError("icp_refine expects float64 init of shape (batch, 4, 4)")
    if source_of_seg is None and source.shape[0] != B:
        raise InvalidArgumentError("icp_refine: without source_of_seg the source batch must equal the target batch")
    for t, name in ((source, "source"), (target, "target"), (init, "init")):
        _capi.require_cuda(t, f"icp_refine({name})")
    source, target, init = source.contiguous(), target.contiguous(), init.contiguous()
    sel = None if source_of_seg is None else _as_i32(source_of_seg, target.device)
    dev = target.device
    T = torch.empty(B, 4, 4, dtype=torch.float64, device=dev)
    fit = torch.empty(B, dtype=torch.float64, device=dev)
    rmse = torch.empty(B, dtype=torch.float64, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    p = _capi.ptr
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().caae_icp_refine(
            B, source.shape[1], source.shape[2], p(source), p(sel), target.shape[1], p(target), p(init), float(radius),
            float(radius_decay), int(outer), int(max_iteration), float(relative_fitness), float(relative_rmse),
            p(T), p(fit), p(rmse), p(iters), _stream(target)), "caae_icp_refine")
    return T, fit, rmse, iters


# ---- ADD / ADD-S ------------------------------------------------------------------------------------------

def add_metrics(source: torch.Tensor, T_gt: torch.Tensor, T_pred: torch.Tensor, source_of_seg=None):
    """Pose errors of a batch of segments behind the prediction (evaluate…:571-575) or the ICP refinement
    (:615-624): ADD = mean_x |(R x + t) - (R^ x + t^)| and ADD-S = mean_x min_y |(R x + t) - (R^ y + t^)| over the model
    points x.  source f32[nsrc,ns,>=3] (the object models; `source_of_seg` i32[B] picks the model of each segment),
    T_gt / T_pred f64[B,4,4].  Returns (add f64[B], add_s f64[B]) in the unit of the models (metres).  The nearest
    neighbour search of ADD-S is the chamfer kernel (nn_distance)."""
    if source.dim() != 3 or source.shape[2] < 3 or source.dtype != torch.float32:
        raise InvalidArgumentError("add_metrics expects float32 source of shape (models, points, >=3)")
    B = T_gt.shape[0]
    for T in (T_gt, T_pred):
        if T.shape != (B, 4, 4) or T.dtype != torch.float64:
            raise InvalidArgumentError("add_metrics expects float64 poses of shape (batch, 4, 4)")
    if source_of_seg is None and source.shape[0] != B:
        raise InvalidArgumentError("add_metrics: without source_of_seg the source batch must equal the pose batch")
    for t, name in ((source, "source"), (T_gt, "T_gt"), (T_pred, "T_pred")):
        _capi.require_cuda(t, f"add_metrics({name})")
    source, T_gt, T_pred = source.contiguous(), T_gt.contiguous(), T_pred.contiguous()
    dev, n = source.device, source.shape[1]
    sel = None if source_of_seg is None else _as_i32(source_of_seg, dev)
    gt = torch.empty(B, n, 3, dtype=torch.float32, device=dev); pred = torch.empty_like(gt)
    d1 = torch.empty(B, n, dtype=torch.float32, device=dev); d2 = torch.empty_like(d1)
    i1 = torch.empty(B, n, dtype=torch.int32, device=dev); i2 = torch.empty_like(i1)
    add = torch.empty(B, dtype=torch.float64, device=dev); adds = torch.empty_like(add)
    p, lib = _capi.ptr, _capi.lib()
    with torch.cuda.device(dev):
        st = _stream(source)
        _capi.check(lib.caae_pose_transform_models(B, n, source.shape[2], p(source), p(sel), p(T_gt), p(T_pred), p(gt),
                                                   p(pred), st), "caae_pose_transform_models")
        _capi.check(lib.caae_nn_distance(B, n, p(gt), n, p(pred), p(d1), p(i1), p(d2), p(i2), st), "caae_nn_distance")
        _capi.check(lib.caae_add_reduce(B, n, p(gt), p(pred), p(d1), p(add), p(adds), st), "caae_add_reduce")
    return add, adds
