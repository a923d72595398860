"""Spherical occluders — drop-in for ``get_random_spherical_occluder`` of the reference's
``utils/generate_occluder.py`` (:38-81), batched, on the GPU."""
from __future__ import annotations

import torch

from .._capi import InvalidArgumentError
from ..synthesis import SegmentSynthesizer


def get_random_spherical_occluder(translation: torch.Tensor, dataset: str = "ycbv", z_centers=None, z_points=None):
    """translation f32[B,3] (CUDA).  Returns occluder f32[B,400,3]: two blobs of 200 points ~ N(centre, 0.01),
    centres x ~ N(0, Wnear/10), y ~ N(0, Hnear/10), z ~ N((near + t_z)/2, (t_z - near)/6); rows alternate
    blob 1 / blob 2 as in the reference.  z_centers f32[B,2,3] / z_points f32[B,2,200,3] are the standard
    normal draws (generated on the device when omitted)."""
    if dataset != "ycbv":
        raise InvalidArgumentError("only the 'ycbv' camera is reachable from train_cloudAAE_ycbv.py")
    b = translation.shape[0]
    syn = SegmentSynthesizer.for_occluder_only(translation.device)
    return syn.occluder(translation, z_centers, z_points)
