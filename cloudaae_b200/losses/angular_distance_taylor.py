"""Rotation error on SO(3) — drop-in for the reference's ``losses/angular_distance_taylor.py``
(skew_symmetric :6-27, exponential_map :30-66, get_rotation_error :102-116), float64 like the reference.

``get_rotation_error`` runs on the hand-written kernel (``caae_pose_losses``: Rodrigues formula with
the theta^2 < 1e-2 Taylor guard, theta = acos(clip((tr(R_l R_p^T) - 1)/2, +-0.9999999)), exact
forward-mode gradient).  ``exponential_map`` / ``skew_symmetric`` are small host-side helpers used by
the synthesis code to turn dataset axis-angles into rotation matrices."""
from __future__ import annotations

import torch

from ._pose import pose_errors


def skew_symmetric(axag_unit: torch.Tensor) -> torch.Tensor:
    """(B,3) -> (B,3,3) cross-product matrices."""
    z = torch.zeros_like(axag_unit[:, 0])
    x, y, w = axag_unit[:, 0], axag_unit[:, 1], axag_unit[:, 2]
    return torch.stack([torch.stack([z, -w, y], 1), torch.stack([w, z, -x], 1), torch.stack([-y, x, z], 1)], 1)


def exponential_map(axag: torch.Tensor, EPS: float = 1e-2) -> torch.Tensor:
    """Rodrigues' formula, Taylor expansion for theta^2 < EPS.  (B,3) -> (B,3,3), dtype of the input
    (the reference feeds float64)."""
    ss = skew_symmetric(axag)
    theta_sq = (axag * axag).sum(dim=1)
    small = theta_sq < EPS
    safe = torch.where(small, torch.ones_like(theta_sq), theta_sq)
    theta = torch.sqrt(safe)
    t4, t6, t8 = theta_sq * theta_sq, theta_sq ** 3, theta_sq ** 4
    term_1 = torch.where(small, 1 - (theta_sq / 6) + (t4 / 120) - (t6 / 5040) + (t8 / 362880), torch.sin(theta) / theta)
    term_2 = torch.where(small, 0.5 - (theta_sq / 24) + (t4 / 720) - (t6 / 40320) + (t8 / 3628800),
                         (1 - torch.cos(theta)) / safe)
    eye = torch.eye(3, dtype=axag.dtype, device=axag.device).unsqueeze(0)
    return eye + term_1[:, None, None] * ss + term_2[:, None, None] * torch.matmul(ss, ss)


def get_rotation_error(pred: torch.Tensor, label: torch.Tensor):
    """Return (mean, per-sample) angular distance in SO(3) between axis-angle pred and label (B,3);
    float64 results, differentiable w.r.t. pred."""
    zeros = torch.zeros(pred.shape[0], 3, dtype=torch.float32, device=pred.device)
    per, _ = pose_errors(pred, label, zeros, zeros)
    return per.mean(), per
