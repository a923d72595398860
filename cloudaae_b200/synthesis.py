"""Fused on-line segment synthesis for the training step (train_cloudAAE_ycbv.py:96-117 + :206-217).

Per batch, on the current CUDA stream, graph-capturable:
   Philox draws -> pose transform + occluder + both spherical flips (1 kernel) -> hidden point removal
   + visible-prefix selection for the occluded cloud (first num_point visible points = network input)
   and for the bare object (first 4*num_point visible points = chamfer target) -> sensor noise draws.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _capi
from .utils.sample_pose_in_frustum import YCBV, get_frustum

NUM_MODEL_POINTS = 2048
NUM_OCCLUDER_POINTS = 400
HPR_PARAM = np.float32(0.8 * np.pi)
NOISE_STD = 0.004 / 3.0  # train_cloudAAE_ycbv.py:216


class SegmentSynthesizer:
    def __init__(self, models_xyz: torch.Tensor, batch_size: int, num_point: int = 256, seed: int = 0):
        """models_xyz f32[num_class, 2048, 3] on the GPU (xyz columns of obj_models.tfrecords)."""
        assert models_xyz.is_cuda and models_xyz.dtype == torch.float32 and models_xyz.dim() == 3
        self.models = models_xyz.contiguous()
        self.dev = models_xyz.device
        self.B, self.N = batch_size, num_point
        self.nm, self.no = models_xyz.shape[1], NUM_OCCLUDER_POINTS
        self.seed = seed
        self.lib = _capi.lib()
        _, hnear, wnear, _, _ = get_frustum(**YCBV)
        self.hnear, self.wnear, self.near = float(hnear), float(wnear), float(YCBV["nearDist"])
        self.flip_pow = float(np.power(np.float32(10.0), HPR_PARAM))
        f32 = dict(dtype=torch.float32, device=self.dev)
        B, n = batch_size, self.nm + self.no
        self.points = torch.empty(B, n, 3, **f32)
        self.flip_all = torch.empty(B, n, 3, **f32)
        self.flip_org = torch.empty(B, self.nm, 3, **f32)
        self.z_centers = torch.zeros(B, 2, 3, **f32)
        self.z_points = torch.zeros(B, 2, self.no // 2, 3, **f32)
        self.pad_u = torch.zeros(B, num_point, **f32)       # defined even when synthesize(draw=False) runs first
        self.pad_u_org = torch.zeros(B, 4 * num_point, **f32)
        # the three products of a batch live in ONE flat buffer so a consumer can take a snapshot with one copy
        self.out_flat = torch.zeros(B * 6 * num_point * 3, **f32)
        nv = B * num_point * 3
        self.visible = self.out_flat[:nv].view(B, num_point, 3)
        self.target = self.out_flat[nv:5 * nv].view(B, 4 * num_point, 3)
        self.noise = self.out_flat[5 * nv:].view(B, num_point, 3)
        self.num_vis = torch.empty(B, dtype=torch.int32, device=self.dev)
        self.num_vis_org = torch.empty(B, dtype=torch.int32, device=self.dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=self.dev)  # bumped once per batch

    _OCC = {}

    @classmethod
    def for_occluder_only(cls, device):
        key = str(device)
        if key not in cls._OCC:
            cls._OCC[key] = cls(torch.zeros(1, NUM_MODEL_POINTS, 3, device=device), 1)
        return cls._OCC[key]

    def _c(self, name, *args):
        _capi.check(getattr(self.lib, name)(*args, torch.cuda.current_stream(self.dev).cuda_stream), name)

    def _fill(self, t, stream_id, uniform=0):
        self._c("caae_philox_fill", t.numel(), t.data_ptr(), self.seed, stream_id, self.counter.data_ptr(), uniform)

    def draw(self):
        """Fresh random draws for one batch (device-side Philox; the counter advances per batch)."""
        self.counter.add_(1)
        self._fill(self.z_centers, 1); self._fill(self.z_points, 2); self._fill(self.noise, 3)
        self._fill(self.pad_u, 4, 1); self._fill(self.pad_u_org, 5, 1)
        self.noise.mul_(NOISE_STD)

    def occluder(self, translation, z_centers=None, z_points=None):
        b = translation.shape[0]
        f32 = dict(dtype=torch.float32, device=self.dev)
        zc = torch.randn(b, 2, 3, **f32) if z_centers is None else z_centers.contiguous().float()
        zp = torch.randn(b, 2, self.no // 2, 3, **f32) if z_points is None else z_points.contiguous().float()
        pts = torch.empty(b, self.nm + self.no, 3, **f32)
        fa = torch.empty_like(pts); fo = torch.empty(b, self.nm, 3, **f32)
        models = torch.zeros(1, self.nm, 3, **f32)
        cls = torch.zeros(b, dtype=torch.int32, device=self.dev)
        ax = torch.zeros(b, 3, **f32)
        self._c("caae_synth_points", b, self.nm, self.no, models.data_ptr(), cls.data_ptr(), ax.data_ptr(),
                translation.contiguous().data_ptr(), zc.data_ptr(), zp.data_ptr(), self.hnear, self.wnear, self.near,
                self.flip_pow, pts.data_ptr(), fa.data_ptr(), fo.data_ptr())
        return pts[:, self.nm:].contiguous()

    def synthesize(self, class_id, axisangle, translation, draw: bool = True):
        """class_id i32[B], axisangle / translation f32[B,3] (pose records).  Returns
        (visible f32[B,N,3], target f32[B,4N,3], noise f32[B,N,3]) — views of reused buffers."""
        B, n = self.B, self.nm + self.no
        if draw:
            self.draw()
        p = _capi.ptr
        self._c("caae_synth_points", B, self.nm, self.no, p(self.models), p(class_id), p(axisangle), p(translation),
                p(self.z_centers), p(self.z_points), self.hnear, self.wnear, self.near, self.flip_pow, p(self.points),
                p(self.flip_all), p(self.flip_org))
        # both hidden-point-removal problems (occluded cloud -> network input, bare object -> chamfer target)
        # in one launch of 2B CTAs
        self._c("caae_hpr_select_pair", B, n, p(self.flip_all), self.N, p(self.pad_u), p(self.visible), p(self.num_vis),
                self.nm, p(self.flip_org), 4 * self.N, p(self.pad_u_org), p(self.target), p(self.num_vis_org),
                p(self.points), n)
        return self.visible, self.target, self.noise


def load_models_xyz(path: str | None = None, device="cuda") -> torch.Tensor:
    """The 21 YCB object models (xyz).  `path` may be the reference's obj_models.tfrecords or an .npy
    [21,2048,3]; defaults to the fixture committed with the tests."""
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "ycb_models_xyz.npy")
    if path.endswith(".npy"):
        arr = np.load(path)
    else:
        from .data.tfrecord import read_object_models
        arr = read_object_models(path)[:, :, :3]
    return torch.from_numpy(np.ascontiguousarray(arr, np.float32)).to(device)
