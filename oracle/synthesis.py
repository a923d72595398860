"""NumPy / SciPy restatement of the reference's on-line segment synthesis — TEST INFRASTRUCTURE ONLY.

Follows train_cloudAAE_ycbv.py:57-117, 206-226, utils/hidden_point_removal.py:6-73,
utils/generate_occluder.py:38-81, utils/sample_pose_in_frustum.py:42-70 and
losses/angular_distance_taylor.py:6-66.  TensorFlow's RNG cannot be reproduced, so every random
draw is an explicit argument (SURVEY.md §7 hard part 7).  PINNED against the reference's own utilities executed in
place (oracle/ref_py; tests/test_ref_py_pins_synthesis_oracle.py: occluder bit-exact, flips to a few fp32 ulps, visible
sets identical through the same scipy.spatial.ConvexHull call) and against the vectors committed from that run
(tests/golden/ref_py_synth_golden.npz).  SciPy's version is the one thing the reference does not pin.
"""
from __future__ import annotations

import math

import numpy as np

# get_frustum(vertical_fov=45., nearDist=0.5, ratio=58/45): the degrees are fed to tan() as radians
# (utils/generate_occluder.py:48-57, utils/sample_pose_in_frustum.py:45-46) — replicate, don't fix.
HNEAR_YCBV = np.float32(2.0) * np.tan(np.float32(45.0) / np.float32(2.0), dtype=np.float32) * np.float32(0.5)
WNEAR_YCBV = np.float32(HNEAR_YCBV * np.float32(58.0 / 45.0))
NEAR_DIST_YCBV = np.float32(0.5)
HPR_PARAM = np.float32(0.8 * math.pi)  # train_cloudAAE_ycbv.py:105


def skew_symmetric(axag: np.ndarray) -> np.ndarray:
    """losses/angular_distance_taylor.py:6-27. axag f64[B,3] -> f64[B,3,3]."""
    z = np.zeros(axag.shape[0], np.float64)
    x, y, w = axag[:, 0], axag[:, 1], axag[:, 2]
    return np.stack([np.stack([z, -w, y], 1), np.stack([w, z, -x], 1), np.stack([-y, x, z], 1)], 1)


def exponential_map(axag: np.ndarray, eps: float = 1e-2) -> np.ndarray:
    """Rodrigues formula with the reference's Taylor guard (angular_distance_taylor.py:30-66). float64."""
    axag = np.asarray(axag, np.float64)
    ss = skew_symmetric(axag)
    theta_sq = np.sum(np.square(axag), axis=1)
    small = theta_sq < eps
    theta = np.sqrt(theta_sq)
    t4 = theta_sq * theta_sq
    t6 = theta_sq * theta_sq * theta_sq
    t8 = theta_sq * theta_sq * theta_sq * theta_sq
    with np.errstate(divide="ignore", invalid="ignore"):
        term1 = np.where(small, 1 - (theta_sq / 6) + (t4 / 120) - (t6 / 5040) + (t8 / 362880), np.sin(theta) / theta)
        term2 = np.where(small, 0.5 - (theta_sq / 24) + (t4 / 720) - (t6 / 40320) + (t8 / 3628800),
                         (1 - np.cos(theta)) / theta_sq)
    eye = np.eye(3, dtype=np.float64)[None]
    return eye + term1[:, None, None] * ss + term2[:, None, None] * np.matmul(ss, ss)


def rotation_error(pred_axag: np.ndarray, label_axag: np.ndarray):
    """get_rotation_error (angular_distance_taylor.py:102-116): geodesic angle, float64."""
    rp, rl = exponential_map(pred_axag), exponential_map(label_axag)
    r = np.matmul(rl, np.transpose(rp, (0, 2, 1)))
    tr = (np.trace(r, axis1=1, axis2=2) - 1) / 2
    theta = np.arccos(np.clip(tr, -0.9999999, 0.9999999))
    return theta.mean(), theta


def transform_object_model(model_xyz: np.ndarray, axisangle: np.ndarray, translation: np.ndarray) -> np.ndarray:
    """get_rotation_matrix + transform_object_model (train_cloudAAE_ycbv.py:79-93).
    model_xyz f32[B,2048,3]; R = f32(expmap_f64(axag)); P = M @ R^T + t in float32."""
    r = exponential_map(np.asarray(axisangle, np.float32).astype(np.float64)).astype(np.float32)
    rot = np.matmul(np.asarray(model_xyz, np.float32), np.transpose(r, (0, 2, 1)))
    return (rot + np.asarray(translation, np.float32)[:, None, :]).astype(np.float32)


def spherical_occluder(trans_z: np.ndarray, z_centers: np.ndarray, z_points: np.ndarray) -> np.ndarray:
    """get_random_spherical_occluder(x, 'ycbv') (utils/generate_occluder.py:38-81) with the normal
    draws made explicit.  trans_z f32[B]; z_centers f32[B,2,3] and z_points f32[B,2,200,3] are
    standard-normal draws.  Returns f32[B,400,3] with the two blobs interleaved (:76-79)."""
    tz = np.asarray(trans_z, np.float32)
    zc = np.asarray(z_centers, np.float32)
    zp = np.asarray(z_points, np.float32)
    cx = zc[:, :, 0] * np.float32(WNEAR_YCBV / np.float32(10.0))
    cy = zc[:, :, 1] * np.float32(HNEAR_YCBV / np.float32(10.0))
    mean_z = (NEAR_DIST_YCBV + tz) / np.float32(2.0)
    std_z = (tz - NEAR_DIST_YCBV) / np.float32(6.0)
    cz = zc[:, :, 2] * std_z[:, None] + mean_z[:, None]
    centers = np.stack([cx, cy, cz], -1).astype(np.float32)  # [B,2,3]
    pts = zp * np.float32(0.01) + centers[:, :, None, :]     # [B,2,200,3]
    # concat([x1,y1,z1,x2,y2,z2], axis=1) then reshape(-1,3): row i -> blob1[i], blob2[i]
    return np.transpose(pts, (0, 2, 1, 3)).reshape(pts.shape[0], 400, 3).astype(np.float32)


def spherical_flip(points: np.ndarray, param: np.float32 = HPR_PARAM):
    """sphericalFlip / sphericalFlip_org with center = 0 (utils/hidden_point_removal.py:6-24, 51-68).
    points f32[B,P,3] -> (flipped f32[B,P+1,3], org f32[B,P+1,3]), a zero row (the viewpoint) appended."""
    p = np.asarray(points, np.float32)
    norm = np.sqrt(np.sum(p * p, axis=2, dtype=np.float32), dtype=np.float32)
    big_r = (norm.max(axis=1, keepdims=True) * np.power(np.float32(10.0), np.float32(param))).astype(np.float32)
    tmp = (np.float32(2.0) * (big_r - norm))[:, :, None] * p
    flipped = (tmp / norm[:, :, None] + p).astype(np.float32)
    zero = np.zeros((p.shape[0], 1, 3), np.float32)
    return np.concatenate([flipped, zero], 1), np.concatenate([p, zero], 1)


def convex_hull_visible(flipped: np.ndarray, org: np.ndarray, pad_draws=None):
    """convexHull (utils/hidden_point_removal.py:27-48).  Returns (visiblePoints f32[B,P+1,3],
    num_vis i64[B], visible_ids list).  pad_draws: optional list of uniform [0,1) arrays used to pick
    the padding indices (np.random.choice with replacement in the reference); default pads by
    cycling through the visible ids so the result is deterministic."""
    from scipy.spatial import ConvexHull
    b, p1, _ = flipped.shape
    vis = np.zeros((b, p1, 3), np.float32)
    num = np.zeros(b, np.int64)
    ids_out = []
    for k in range(b):
        flag = np.zeros(p1, int)
        hull = ConvexHull(flipped[k])
        flag[hull.vertices[:-1]] = 1
        ids = np.where(flag == 1)[0][:-1]
        npad = p1 - len(ids)
        if pad_draws is None:
            pad = ids[np.arange(npad) % max(len(ids), 1)] if len(ids) else np.zeros(npad, int)
        else:
            pad = ids[np.minimum((np.asarray(pad_draws[k])[:npad] * len(ids)).astype(int), len(ids) - 1)]
        vis[k] = org[k, np.concatenate((ids, pad))]
        num[k] = len(ids)
        ids_out.append(ids)
    return vis, num, ids_out
