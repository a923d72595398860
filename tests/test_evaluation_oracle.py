"""CPU checks of oracle/evaluation.py (the checker of the evaluation front end and ICP): the restated
open3d / NumPy algorithms against independent library computations (SciPy KD-tree, SciPy rotations)."""
import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

import cases
from oracle import evaluation as E


def _frame(record=0, classes=(0, 3, 7), splat=1, seed=0):
    clouds = cases.posed_ycb_clouds(record)
    return E.render_frame(clouds[list(classes)], list(classes), splat=splat, seed=seed)


def test_get_pointcloud_is_the_pinhole_back_projection():
    depth, _ = _frame()
    fx, fy, cx, cy, fac = [float(v) for v in E.YCBV_INTRINSICS]
    xyz = E.get_pointcloud(depth, fx, fy, cx, cy, fac).reshape(480, 640, 3)
    assert xyz.dtype == np.float32
    v, u = 100, 517
    z = depth[v, u] / fac
    np.testing.assert_allclose(xyz[v, u], [(u - cx) * z / fx, (v - cy) * z / fy, z], rtol=1e-6)
    assert (xyz[depth == 0] == 0).all()


def test_segment_extract_masks_and_filter():
    depth, label = _frame()
    for c in (0, 3, 7):
        org, flt, pix, mean = E.segment_extract(depth, label, E.YCBV_INTRINSICS, c, 0.2)
        assert org.shape[0] == int(((label == c + 1) & (depth != 0)).sum())
        assert 100 < flt.shape[0] < org.shape[0]            # the stray pixels are farther than 0.2 m
        assert (np.linalg.norm(flt - mean, axis=1) <= 0.2 + 1e-6).all()
        assert (np.diff(pix) > 0).all() and (label.reshape(-1)[pix] == c + 1).all()
    org, flt, pix, mean = E.segment_extract(depth, label, E.YCBV_INTRINSICS, 11, 0.2)  # class not in the frame
    assert org.shape[0] == 0 and flt.shape[0] == 0 and np.isnan(mean).all()


def test_radius_counts_equal_kdtree_ball_query():
    depth, label = _frame()
    _, flt, _, _ = E.segment_extract(depth, label, E.YCBV_INTRINSICS, 7, 0.2)
    cnt = E.radius_neighbour_counts(flt, 0.02)
    assert (cnt == E.radius_neighbour_counts(flt, 0.02, brute=True)).all()
    g = np.arange(8, dtype=np.float32) * np.float32(0.005)   # lattice: pairs at distance == radius up to rounding
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    assert (E.radius_neighbour_counts(lat, 0.02) == E.radius_neighbour_counts(lat, 0.02, brute=True)).all()
    tree = cKDTree(flt.astype(np.float64))
    ball = np.array([len(x) for x in tree.query_ball_point(flt.astype(np.float64), 0.02)])
    assert np.abs(ball - cnt).max() <= 1          # the tree uses <=, open3d/FLANN use <; boundary pairs only
    idx = E.get_outlier_idx(flt)
    assert (cnt[idx] > 100).all() and len(idx) >= 512
    few = flt[:300]
    assert (E.get_outlier_idx(few) == np.arange(300)).all()   # fewer than 512 inliers -> everything is kept


def test_fps_random_is_farthest_point_sampling():
    rng = np.random.default_rng(3)
    pts = rng.standard_normal((700, 3)).astype(np.float32)
    idx = E.FPS_random(pts, 64, first_idx=17)
    assert idx[0] == 17 and len(set(idx.tolist())) == 64
    d = np.full(700, np.inf)
    for i in range(63):
        d = np.minimum(d, ((pts[idx[i]].astype(np.float64) - pts) ** 2).sum(1))
        assert idx[i + 1] == int(np.argmax(d))
    # more samples than points: the tail repeats the first maximum (index of an all-zero distance array = 0)
    tail = E.FPS_random(pts[:5], 9, first_idx=2)
    assert set(tail[:5].tolist()) == {0, 1, 2, 3, 4} and (tail[5:] == 0).all()


def test_umeyama_recovers_a_rigid_motion():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((200, 3))
    R = Rotation.from_rotvec([0.3, -0.2, 0.5]).as_matrix()
    t = np.array([0.1, 0.2, -0.3])
    T = E.umeyama_rigid(x, x @ R.T + t)
    np.testing.assert_allclose(T[:3, :3], R, atol=1e-12)
    np.testing.assert_allclose(T[:3, 3], t, atol=1e-12)
    assert (E.umeyama_rigid(x[:0], x[:0]) == np.eye(4)).all()


def test_icp_refine_pulls_a_perturbed_pose_back():
    models = cases.ycb_models()
    t, a, c = cases.ycb_poses()
    per = len(c) // 21
    for cls in (1, 9):
        rec = cls * per
        R = Rotation.from_rotvec(a[rec].astype(np.float64)).as_matrix()
        posed = models[cls].astype(np.float64) @ R.T + t[rec]
        target = posed[::8].astype(np.float32)              # 256 points of the posed model
        dR = Rotation.from_rotvec([0.02, -0.03, 0.025]).as_matrix()
        init = np.eye(4)
        init[:3, :3] = dR @ R
        init[:3, 3] = t[rec] + np.array([0.003, -0.002, 0.004])
        T, fit, rmse, iters = E.icp_refine(models[cls], target, init)
        before = np.abs(models[cls] @ init[:3, :3].T + init[:3, 3] - posed).max()
        after = np.abs(models[cls] @ T[:3, :3].T + T[:3, 3] - posed).max()
        assert after < 0.2 * before and fit > 0.05 and rmse < 0.005 and iters >= 10
        np.testing.assert_allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)


def test_oracle_reproduces_the_committed_golden_vectors():
    """tests/golden/eval_golden.npz (written by tests/golden/make_golden_eval.py) pins the oracle itself: a change
    of NumPy / SciPy or of the restatement that moves any of these outputs must be noticed."""
    import os
    g = np.load(os.path.join(cases.GOLDEN, "eval_golden.npz"))
    depth, label = cases.eval_golden_frame()
    assert (g["frame_checksum"] == [int(depth.astype(np.int64).sum()), int(label.astype(np.int64).sum())]).all()
    for c in cases.EVAL_GOLDEN_CLASSES:
        org, flt, pix, mean = E.segment_extract(depth, label, E.YCBV_INTRINSICS, c, 0.2)
        idx = E.get_outlier_idx(flt)
        assert (g[f"seg{c}_counts"] == [len(org), len(flt), len(idx), E.num_valid_points(idx)]).all()
        assert (g[f"seg{c}_mean"] == mean).all()
        assert (g[f"seg{c}_pix_head"] == pix[:64]).all() and (g[f"seg{c}_inlier_head"] == idx[:64]).all()
        assert (g[f"seg{c}_fps"] == E.FPS_random(flt, 64, 5)).all()
    model, target, init = cases.eval_golden_icp_case()
    T, fit, rmse, it = E.icp_refine(model, target, init)
    np.testing.assert_allclose(T, g["icp_T"], atol=1e-9)
    np.testing.assert_allclose([fit, rmse, it], g["icp_stats"], atol=1e-9)
