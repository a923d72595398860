"""Frustum constants — drop-in for ``get_frustum`` of the reference's ``utils/sample_pose_in_frustum.py``
(:42-70), the only function of that file reachable from train_cloudAAE_ycbv.py."""
from __future__ import annotations

import numpy as np


def get_frustum(vertical_fov, nearDist, farDist, ratio):
    """Hnear = 2*tan(vertical_fov/2)*nearDist etc.  NOTE: like the reference, `vertical_fov` goes to
    tan() as given — the callers pass 45. (degrees) and the reference evaluates tan(22.5 rad); the
    resulting constants are what shaped the training data, so they are reproduced, not fixed.
    Returns (frustum_corners [3,8], Hnear, Wnear, Hfar, Wfar) as float32."""
    f32 = np.float32
    t = np.tan(f32(vertical_fov) / f32(2), dtype=np.float32)
    Hnear = f32(2) * t * f32(nearDist); Wnear = Hnear * f32(ratio)
    Hfar = f32(2) * t * f32(farDist); Wfar = Hfar * f32(ratio)
    up, right, cam = np.array([0, 1, 0], f32), np.array([1, 0, 0], f32), np.array([0, 0, 1], f32)
    fc, nc = cam * f32(farDist), cam * f32(nearDist)
    corners = [fc + up * Hfar / 2 - right * Wfar / 2, fc + up * Hfar / 2 + right * Wfar / 2,
               fc - up * Hfar / 2 - right * Wfar / 2, fc - up * Hfar / 2 + right * Wfar / 2,
               nc + up * Hnear / 2 - right * Wnear / 2, nc + up * Hnear / 2 + right * Wnear / 2,
               nc - up * Hnear / 2 - right * Wnear / 2, nc - up * Hnear / 2 + right * Wnear / 2]
    return np.stack(corners, axis=1).astype(f32), f32(Hnear), f32(Wnear), f32(Hfar), f32(Wfar)


# the 'ycbv' camera of get_random_spherical_occluder (utils/generate_occluder.py:47-51)
YCBV = dict(vertical_fov=45.0, nearDist=0.5, farDist=1.0, ratio=58.0 / 45.0)
