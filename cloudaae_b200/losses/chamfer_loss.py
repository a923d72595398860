"""Chamfer loss — drop-in for the reference's ``losses/chamfer_loss.py`` (get_loss :8-14)."""
from __future__ import annotations

import torch

from .._capi import InvalidArgumentError
from ..tf_ops.nn_distance.tf_nndistance import nn_distance


def get_loss(pred: torch.Tensor, label: torch.Tensor):
    """pred: BxNx3, label: BxNx3.  loss_per_sample = dists_forward + dists_backward (both [B,N], which
    is why the reference only works for equal point counts), loss = mean.  Differentiable w.r.t. both."""
    dists_forward, _, dists_backward, _ = nn_distance(pred, label)
    if dists_forward.shape != dists_backward.shape:
        raise InvalidArgumentError("chamfer_loss.get_loss adds dist1 [B,n] and dist2 [B,m]: n must equal m")
    loss_per_sample = dists_forward + dists_backward
    loss = loss_per_sample.mean()
    return loss, loss_per_sample
