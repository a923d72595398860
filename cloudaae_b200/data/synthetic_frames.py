"""Synthetic YCB-Video-shaped depth / label frames (NumPy, data generation only — no compute path).

The reference tree ships no real frame (`ycb_video_data_tfRecords/` holds pose records only), so tests and
bench.py render the posed object models into frames of the layout evaluate_cloudAAE_ycbv.py:125-160 decodes:
uint16 depth (metres * factor_depth, 0 = invalid), uint8 one-based class labels, 480 x 640.
"""
from __future__ import annotations

import numpy as np

# YCB-Video camera (fx, fy, cx, cy, factor_depth); the reference reads these per record (evaluate…:140-144)
YCBV_INTRINSICS = np.array([1066.778, 1067.487, 312.9869, 241.3109, 10000.0], np.float32)


def render_frame(clouds, class_ids, h=480, w=640, intrinsics=YCBV_INTRINSICS, splat=2, seed=0, n_stray=300):
    """Project posed object clouds (camera frame, metres) into a u16 depth image and a u8 label image
    (one-based labels, nearest surface wins), each point splatted over (2*splat+1)^2 pixels; every pixel
    gets a background depth; `n_stray` isolated pixels per object carry its label at a wrong depth so both
    filters have something to remove."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, factor = [float(v) for v in intrinsics]
    zbuf = np.full((h, w), np.inf)
    label = np.zeros((h, w), np.uint8)
    for pts, c in zip(clouds, class_ids):
        u = np.rint(pts[:, 0] * fx / pts[:, 2] + cx).astype(int)
        v = np.rint(pts[:, 1] * fy / pts[:, 2] + cy).astype(int)
        tmp = np.full((h, w), np.inf)
        for du in range(-splat, splat + 1):
            for dv in range(-splat, splat + 1):
                uu, vv = u + du, v + dv
                ok = np.flatnonzero((uu >= 0) & (uu < w) & (vv >= 0) & (vv < h))
                ok = ok[np.argsort(-pts[ok, 2], kind="stable")]  # far first: the nearest write lands last
                near = pts[ok, 2] < tmp[vv[ok], uu[ok]]
                ok = ok[near]
                tmp[vv[ok], uu[ok]] = pts[ok, 2]
        m = tmp < zbuf
        zbuf[m] = tmp[m]
        label[m] = c + 1
        ys, xs = rng.integers(0, h, n_stray), rng.integers(0, w, n_stray)
        zbuf[ys, xs] = float(pts[:, 2].mean()) + rng.uniform(-0.6, 0.6, n_stray)
        label[ys, xs] = c + 1
    depth = np.where(np.isfinite(zbuf), zbuf, 2.5)
    depth_u16 = np.clip(np.rint(depth * factor), 0, 65535).astype(np.uint16)
    holes = rng.random((h, w)) < 0.02  # invalid depth readings
    depth_u16[holes] = 0
    return depth_u16, label
