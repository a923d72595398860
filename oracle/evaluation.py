"""NumPy restatement of the evaluation front end and ICP refinement — TEST INFRASTRUCTURE ONLY.

Follows the reference's evaluate_cloudAAE_ycbv.py: get_pointcloud (:164-178), segment_not_empty
(:262-272), segment_mean_distance_filter (:219-223), get_outlier_idx (:250-258), FPS_random (:230-247)
and the ICP loop (:606-624).

Third-party arithmetic that is absent from /root/reference:
  * open3d (README.md:21, no version pinned): `remove_radius_outlier(nb_points, radius)` keeps point i when
    the KD-tree radius search around it (squared distance < radius^2, the point itself included, float64)
    returns MORE than nb_points neighbours; `registration_icp(source, target, max_dist, init,
    TransformationEstimationPointToPoint())` with the default ICPConvergenceCriteria (relative_fitness
    1e-6, relative_rmse 1e-6, max_iteration 30) — restated below from its published algorithm.
  * TensorFlow's reduce_mean order is not reproducible; the mean is accumulated in float64 here.
Parity unpinned: the reference holds no test or golden vector for this path and open3d cannot be
installed here.  The random first index of FPS_random (random.randint) is an explicit argument.
"""
from __future__ import annotations

import numpy as np

# YCB-Video camera of the evaluation records (fx, fy, cx, cy, factor_depth) — shapes only; the
# reference reads them from each tfrecord (evaluate…:140-144)
YCBV_INTRINSICS = np.array([1066.778, 1067.487, 312.9869, 241.3109, 10000.0], np.float32)


def get_pointcloud(depth: np.ndarray, fx, fy, cx, cy, depth_scaling_factor) -> np.ndarray:
    """evaluate…:164-178.  depth u16[h,w] -> f32[h*w,3]; every op rounds to float32 on its own."""
    f = np.float32
    depth_meters = depth.astype(np.float32) / f(depth_scaling_factor)
    h, w = depth_meters.shape
    X, Y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    x = (X - f(cx)) * depth_meters / f(fx)
    y = (Y - f(cy)) * depth_meters / f(fy)
    return np.stack([x, y, depth_meters], axis=2).reshape(h * w, 3)


def segment_extract(depth: np.ndarray, label: np.ndarray, intrinsics, class_id: int, threshold: float):
    """segment_not_empty + the boolean_mask of outlier_removal (evaluate…:262-278) for one (frame, class).
    Returns xyz_org f32[n_org,3], xyz_filt f32[n_filt,3], pix_filt i32[n_filt], mean f32[3]."""
    fx, fy, cx, cy, factor = [np.float32(v) for v in intrinsics]
    xyz = get_pointcloud(depth, fx, fy, cx, cy, factor)
    label_flat = label.reshape(-1).astype(np.int64) - 1
    mask = (label_flat == class_id) & (depth.reshape(-1).astype(np.int64) != 0)
    sel = xyz[mask]
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = (sel.astype(np.float64).sum(axis=0) / np.float64(sel.shape[0])).astype(np.float32)
        diff = xyz - mean
        d = np.sqrt((diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2])
        mask_r = mask & (d <= np.float32(threshold))
    return sel, xyz[mask_r], np.flatnonzero(mask_r).astype(np.int32), mean


def _brute_counts(q: np.ndarray, p: np.ndarray, r2, chunk: int = 512) -> np.ndarray:
    out = np.zeros(q.shape[0], np.int64)
    for s in range(0, q.shape[0], chunk):
        qq = q[s:s + chunk]
        dx = qq[:, None, 0] - p[None, :, 0]
        dy = qq[:, None, 1] - p[None, :, 1]
        dz = qq[:, None, 2] - p[None, :, 2]
        d = (dx * dx + dy * dy) + dz * dz
        out[s:s + chunk] = (d < r2).sum(axis=1)
    return out


def radius_neighbour_counts(xyz: np.ndarray, radius: float, brute: bool = False) -> np.ndarray:
    """Number of points (the point itself included) at float64 squared distance < radius^2, the squared
    distance evaluated as (dx*dx + dy*dy) + dz*dz.  A KD-tree brackets the count with two radii 1e-7 apart;
    only queries with a pair inside that band (rounding could matter there) are counted by brute force."""
    p = np.asarray(xyz, np.float32).astype(np.float64)
    r2 = np.float64(radius) * np.float64(radius)
    if brute or p.shape[0] < 64:
        return _brute_counts(p, p, r2)
    from scipy.spatial import cKDTree
    tree = cKDTree(p)
    lo = tree.query_ball_point(p, radius * (1 - 1e-7), return_length=True)
    hi = tree.query_ball_point(p, radius * (1 + 1e-7), return_length=True)
    out = np.asarray(lo, np.int64)
    amb = np.flatnonzero(lo != hi)
    if len(amb):
        out[amb] = _brute_counts(p[amb], p, r2)
    return out


def get_outlier_idx(xyz: np.ndarray, nb_points: int = 100, radius: float = 0.02, min_keep: int = 512) -> np.ndarray:
    """evaluate…:250-258 with open3d's remove_radius_outlier restated."""
    idx = np.flatnonzero(radius_neighbour_counts(xyz, radius) > nb_points)
    if len(idx) < min_keep:
        idx = np.arange(xyz.shape[0])
    return idx.astype(np.int32)


def num_valid_points(inlier_idx: np.ndarray) -> int:
    """`tf.count_nonzero(x['inlier_idx'])` (evaluate…:279): counts non-zero INDEX VALUES, so a kept point 0
    is not counted — replicate, don't fix."""
    return int(np.count_nonzero(inlier_idx))


def calc_distances(p0, points):
    return ((p0 - points) ** 2).sum(axis=1)


def FPS_random(pts: np.ndarray, K: int, first_idx: int) -> np.ndarray:
    """evaluate…:230-247, literally; `first_idx` replaces random.randint(0, n-1)."""
    farthest_pts = np.zeros((K, 3))
    farthest_pts_idx = np.zeros(K)
    farthest_pts[0] = pts[first_idx]
    farthest_pts_idx[0] = first_idx
    distances = calc_distances(farthest_pts[0, 0:3], pts[:, 0:3])
    for i in range(1, K):
        farthest_pts[i] = pts[np.argmax(distances)]
        farthest_pts_idx[i] = np.argmax(distances)
        distances = np.minimum(distances, calc_distances(farthest_pts[i, 0:3], pts[:, 0:3]))
    return farthest_pts_idx.astype(np.int64)


# ---- ICP ------------------------------------------------------------------------------------------

def _correspondences(pcd: np.ndarray, target: np.ndarray, max_dist: float):
    dx = pcd[:, None, 0] - target[None, :, 0]
    dy = pcd[:, None, 1] - target[None, :, 1]
    dz = pcd[:, None, 2] - target[None, :, 2]
    d = (dx * dx + dy * dy) + dz * dz
    j = d.argmin(axis=1)
    dm = d[np.arange(pcd.shape[0]), j]
    ok = dm < np.float64(max_dist) * np.float64(max_dist)
    src = np.flatnonzero(ok)
    n = len(src)
    fitness = n / pcd.shape[0] if pcd.shape[0] else 0.0
    rmse = float(np.sqrt(dm[ok].sum() / n)) if n else 0.0
    return src, j[ok], fitness, rmse


def umeyama_rigid(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Eigen::umeyama(x, y, with_scaling=false): 4x4 T with y ~ R x + t (SVD with the reflection guard)."""
    T = np.eye(4)
    if x.shape[0] == 0:
        return T
    mx, my = x.mean(axis=0), y.mean(axis=0)
    sigma = (y - my).T @ (x - mx) / x.shape[0]
    U, D, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1
    R = U @ np.diag(S) @ Vt
    T[:3, :3] = R
    T[:3, 3] = my - R @ mx
    return T


def registration_icp(source: np.ndarray, target: np.ndarray, max_dist: float, init: np.ndarray,
                     max_iteration: int = 30, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6):
    """open3d registration_icp, point-to-point.  Returns (T, fitness, inlier_rmse, iterations)."""
    T = np.array(init, np.float64)
    src = np.asarray(source, np.float32).astype(np.float64)
    tgt = np.asarray(target, np.float32).astype(np.float64)
    pcd = src @ T[:3, :3].T + T[:3, 3]
    si, tj, fit, rmse = _correspondences(pcd, tgt, max_dist)
    it = 0
    for it in range(1, max_iteration + 1):
        U = umeyama_rigid(pcd[si], tgt[tj])
        T = U @ T
        pcd = pcd @ U[:3, :3].T + U[:3, 3]
        si, tj, nfit, nrmse = _correspondences(pcd, tgt, max_dist)
        done = abs(fit - nfit) < relative_fitness and abs(rmse - nrmse) < relative_rmse
        fit, rmse = nfit, nrmse
        if done:
            break
    return T, fit, rmse, it


def icp_refine(source, target, init, radius=0.01, radius_decay=0.9, outer=10, max_iteration=30):
    """The loop of evaluate…:615-624."""
    T = np.array(init, np.float64)
    fit = rmse = 0.0
    iters = 0
    for _ in range(outer):
        T, fit, rmse, it = registration_icp(source, target, radius, T, max_iteration)
        iters += it
        radius = radius * radius_decay
    return T, fit, rmse, iters


# synthetic YCB-Video-shaped frames (there is no real frame in the reference tree): the generator is a
# data helper of the package, shared by the tests and bench.py
from cloudaae_b200.data.synthetic_frames import render_frame  # noqa: E402,F401


def add_metrics(model_xyz, T_gt, T_pred):
    """ADD and ADD-S of one segment (Hinterstoisser et al. 2012; Xiang et al. 2018): model f32[n,3], poses f64[4,4].
    Points are rounded to float32 after the transform, as the GPU path stores them; the nearest neighbour of ADD-S
    by SciPy's KD-tree."""
    from scipy.spatial import cKDTree
    x = np.asarray(model_xyz, np.float32).astype(np.float64)
    g = (x @ T_gt[:3, :3].T + T_gt[:3, 3]).astype(np.float32).astype(np.float64)
    p = (x @ T_pred[:3, :3].T + T_pred[:3, 3]).astype(np.float32).astype(np.float64)
    add = np.linalg.norm(g - p, axis=1).mean()
    adds = cKDTree(p).query(g)[0].mean()
    return add, adds
