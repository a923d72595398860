"""Translation error — drop-in for the reference's ``losses/trans_distance.py`` (:4-9)."""
from __future__ import annotations

import torch

from ._pose import pose_errors


def get_translation_error(pred: torch.Tensor, label: torch.Tensor):
    """loss_perSample = ||label - pred||_2 over axis 1, loss = mean.  pred, label: (B,3) float32."""
    zeros = torch.zeros_like(pred)
    _, per = pose_errors(zeros.detach(), zeros.detach(), pred, label)
    return per.mean(), per
