"""Hidden point removal — drop-in for the reference's ``utils/hidden_point_removal.py``
(sphericalFlip :6-24, convexHull :27-48, hidden_point_removal :44-48, and the `_org` twins), batched, on the GPU.

The reference runs ``scipy.spatial.ConvexHull`` per sample inside ``tf.py_func``; here visibility is
decided by the sm_100a kernel behind ``caae_hpr_select`` (see csrc/synthesis.cu for the method)."""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _capi
from .._capi import InvalidArgumentError

HPR_PARAM = 0.8 * math.pi  # train_cloudAAE_ycbv.py:105


def _check(points):
    if points.dim() != 3 or points.shape[2] != 3 or points.dtype != torch.float32:
        raise InvalidArgumentError("points must be float32 (batch, num_points, 3)")
    _capi.require_cuda(points, "hidden_point_removal")


def sphericalFlip(points: torch.Tensor, center: torch.Tensor | None = None, param: float = HPR_PARAM):
    """points f32[B,P,3] -> (flippedPoints, orgPoints), each f32[B,P+1,3] with the viewpoint (a zero
    row) appended.  R = max|p| * 10^param; f = 2(R-|p|) p/|p| + p.  center must be 0 (the camera), as
    in both reference call sites (train_cloudAAE_ycbv.py:103-110)."""
    _check(points)
    if center is not None and bool((center != 0).any()):
        raise InvalidArgumentError("only center = 0 (the camera) is supported, as the reference uses it")
    p = points
    norm = torch.sqrt((p * p).sum(dim=2))
    big_r = norm.max(dim=1, keepdim=True).values * float(np.power(np.float32(10.0), np.float32(param)))
    flipped = (2 * (big_r - norm)).unsqueeze(2) * p / norm.unsqueeze(2) + p
    zero = torch.zeros(p.shape[0], 1, 3, dtype=p.dtype, device=p.device)
    return torch.cat([flipped, zero], 1), torch.cat([p, zero], 1)


sphericalFlip_org = sphericalFlip


def convexHull(points: torch.Tensor, orgPoints: torch.Tensor, pad_uniform: torch.Tensor | None = None,
               return_flags: bool = False):
    """points = flipped f32[B,P+1,3] (last row = viewpoint), orgPoints f32[B,P+1,3].
    Returns (visiblePoints f32[B,P+1,3], num_vis_point i64[B]): rows 0..num_vis-1 are the visible
    original points in ascending index order (the highest visible index dropped, as the reference's
    ``visibleId[:-1]`` does), the rest random repeats of visible points (pad_uniform f32[B,P+1] in [0,1),
    cyclic when omitted)."""
    _check(points); _check(orgPoints)
    b, p1, _ = points.shape
    n = p1 - 1
    fl = points[:, :n].contiguous()
    org = orgPoints.contiguous()
    out = torch.empty(b, p1, 3, dtype=torch.float32, device=points.device)
    num = torch.empty(b, dtype=torch.int32, device=points.device)
    flags = torch.empty(b, n, dtype=torch.uint8, device=points.device) if return_flags else None
    pad = None if pad_uniform is None else pad_uniform.contiguous().float()
    with torch.cuda.device(points.device):
        _capi.check(_capi.lib().caae_hpr_select(b, n, fl.data_ptr(), org.data_ptr(), p1, p1, _capi.ptr(pad),
                                                out.data_ptr(), num.data_ptr(), _capi.ptr(flags),
                                                _capi.stream_of(points)), "caae_hpr_select")
    if return_flags:
        return out, num.long(), flags
    return out, num.long()


def hidden_point_removal(x: dict) -> dict:
    """dict-in / dict-out form of the reference's tf.data map function (:44-48)."""
    x["visiblePoints"], x["num_vis_point"] = convexHull(x["flippedPoints"], x["orgPoints"])
    return x


def hidden_point_removal_org(x: dict) -> dict:
    x["visiblePoints_org"], x["num_vis_point_org"] = convexHull(x["flippedPoints_org"], x["orgPoints_org"])
    return x
