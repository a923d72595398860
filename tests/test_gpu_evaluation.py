"""GPU parity of the evaluation front end and the ICP refinement against
oracle/evaluation.py, through the public functions -> ctypes -> C ABI."""
import random

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

import cases
from oracle import evaluation as E

pytestmark = pytest.mark.gpu

from cloudaae_b200 import evaluate_cloudAAE_ycbv as EV  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def np_(t):
    return t.detach().cpu().numpy()


FRAME_CLASSES = [(0, 3, 7), (1, 9, 14), (20, 5, 12)]


@pytest.fixture(scope="module")
def frames():
    depth, label = [], []
    for f, cl in enumerate(FRAME_CLASSES):
        clouds = cases.posed_ycb_clouds(f)
        d, l = E.render_frame(clouds[list(cl)], list(cl), splat=1 + (f % 2), seed=f)
        depth.append(d); label.append(l)
    depth, label = np.stack(depth), np.stack(label)
    intr = np.tile(E.YCBV_INTRINSICS, (len(FRAME_CLASSES), 1))
    intr[1, 0] *= 1.01          # per-frame intrinsics are really per frame
    intr[2, 4] = 5000.0
    thr = np.full(21, 0.2, np.float32)
    thr[9] = 0.05               # a tight threshold cuts into the object itself
    segs = [(f, c) for f, cl in enumerate(FRAME_CLASSES) for c in cl] + [(0, 11), (2, 0)]  # two absent classes
    return depth, label, intr.astype(np.float32), thr, segs


def _front_end(frames, cap=40000):
    depth, label, intr, thr, segs = frames
    return EV.SegmentFrontEnd(cu(depth), cu(label), cu(intr), cu(thr), cap=cap), segs


def test_get_pointcloud_bit_exact(frames):
    depth, _, intr, _, _ = frames
    got = np_(EV.get_pointcloud(cu(depth[2]), *[float(v) for v in intr[2]]))
    want = E.get_pointcloud(depth[2], *intr[2])
    assert got.shape == want.shape and (got == want).all()


def test_segment_extract_bit_exact(frames):
    depth, label, intr, thr, _ = frames
    fe, segs = _front_end(frames)
    out = fe.extract([s[0] for s in segs], [s[1] for s in segs])
    torch.cuda.synchronize()
    for s, (f, c) in enumerate(segs):
        org, flt, pix, mean = E.segment_extract(depth[f], label[f], intr[f], c, thr[c])
        n_org, n_flt = int(out["n_org"][s]), int(out["num_point_after_filter"][s])
        assert n_org == org.shape[0] and n_flt == flt.shape[0], (s, f, c)
        assert (np_(out["xyz_org"][s, :n_org]) == org).all()
        assert (np_(out["xyz_org_distance_filtered"][s, :n_flt]) == flt).all()
        assert (np_(out["pix"][s, :n_flt]) == pix).all()
        if n_org:
            assert (np_(out["mean"][s]) == mean).all()
        else:
            assert np.isnan(np_(out["mean"][s])).all()


def test_segment_extract_cap_truncates_but_counts_everything(frames):
    depth, label, intr, thr, _ = frames
    fe, segs = _front_end(frames, cap=1000)
    out = fe.extract([0], [0])
    org, flt, pix, _ = E.segment_extract(depth[0], label[0], intr[0], 0, thr[0])
    assert int(out["n_org"][0]) == org.shape[0] > 1000 and int(out["num_point_after_filter"][0]) == flt.shape[0]
    assert (np_(out["xyz_org_distance_filtered"][0]) == flt[:1000]).all() and (np_(out["pix"][0]) == pix[:1000]).all()


def _padded(clouds, cap):
    xyz = np.zeros((len(clouds), cap, 3), np.float32)
    for i, c in enumerate(clouds):
        xyz[i, :len(c)] = c
    return xyz, np.array([len(c) for c in clouds], np.int32)


def test_radius_outliers_equal_oracle(frames):
    depth, label, intr, thr, _ = frames
    clouds = [E.segment_extract(depth[0], label[0], intr[0], c, 0.2)[1] for c in (0, 7)]
    clouds.append(clouds[1][:300])                                   # < 512 inliers -> keep all
    g = np.arange(12, dtype=np.float32) * np.float32(0.005)          # lattice: many pairs at distance == radius
    clouds.append(np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32))
    clouds.append(np.zeros((0, 3), np.float32))                      # empty segment
    cap = max(len(c) for c in clouds) + 5
    xyz, n = _padded(clouds, cap)
    for nb, radius, min_keep in ((100, 0.02, 512), (30, 0.02, 1)):
        idx, n_in = EV._radius_outliers(cu(xyz), cu(n), nb, radius, min_keep)
        for s, c in enumerate(clouds):
            want = E.get_outlier_idx(c, nb, radius, min_keep)
            assert int(n_in[s]) == len(want), (s, nb)
            assert (np_(idx[s, :len(want)]) == want).all()
            assert (np_(idx[s, len(want):]) == 0).all()
    got = np_(EV.get_outlier_idx(cu(clouds[0]), 100, 0.02, 0.5))
    assert (got == E.get_outlier_idx(clouds[0])).all()


def test_fps_random_equals_numpy_reference(frames):
    depth, label, intr, thr, _ = frames
    big = E.segment_extract(depth[0], label[0], intr[0], 0, 0.2)[1]
    rng = np.random.default_rng(1)
    dup = rng.standard_normal((400, 3)).astype(np.float32)
    dup = np.concatenate([dup, dup[rng.permutation(400)]])           # exact ties: np.argmax takes the first
    clouds = [big, big[:2049], dup, big[:100], big[:1]]
    firsts = [5, 2048, 799, 99, 0]
    cap = len(big)
    xyz, n = _padded(clouds, cap)
    K = 256
    idx, pts = EV._fps_seeded(cu(xyz), cu(n), cu(np.array(firsts, np.int32)), K)
    for s, c in enumerate(clouds):
        want = E.FPS_random(c, K, firsts[s])
        assert (np_(idx[s]) == want).all(), s
        assert (np_(pts[s]) == c[want]).all()
    one = np_(EV.FPS_random(cu(dup), 64, first_idx=3))
    assert (one == E.FPS_random(dup, 64, 3)).all()


def test_front_end_chain_equals_oracle_chain(frames):
    depth, label, intr, thr, _ = frames
    fe, segs = _front_end(frames)
    out = fe.run([s[0] for s in segs], [s[1] for s in segs], 256, rng=random.Random(7))
    rng = random.Random(7)
    chain = []
    for f, c in segs:
        org, flt, pix, mean = E.segment_extract(depth[f], label[f], intr[f], c, thr[c])
        idx = E.get_outlier_idx(flt) if len(flt) else np.zeros(0, np.int32)
        chain.append((flt, idx))
    first_in = [rng.randint(0, max(len(i) - 1, 0)) for _, i in chain]
    first_org = [rng.randint(0, max(len(f) - 1, 0)) for f, _ in chain]
    keep = np_(out["keep"])
    assert keep[:9].all() and not keep[9:].any()
    for s, (flt, idx) in enumerate(chain):
        assert int(out["n_inlier"][s]) == len(idx)
        assert int(out["num_valid_points_in_segment"][s]) == E.num_valid_points(idx)
        if len(flt) == 0:
            continue
        inl = flt[idx]
        assert (np_(out["xyz_inlier_full"][s, :len(idx)]) == inl).all()
        assert (np_(out["xyz_inlier"][s]) == inl[E.FPS_random(inl, 256, first_in[s])]).all()
        assert (np_(out["xyz"][s]) == flt[E.FPS_random(flt, 256, first_org[s])]).all()


def _icp_case(cls, record_offset, seed):
    models = cases.ycb_models()
    t, a, c = cases.ycb_poses()
    per = len(c) // 21
    rec = cls * per + record_offset
    rng = np.random.default_rng(seed)
    R = Rotation.from_rotvec(a[rec].astype(np.float64)).as_matrix()
    posed = models[cls].astype(np.float64) @ R.T + t[rec]
    vis = posed[posed[:, 2] < np.median(posed[:, 2])]                # the camera-facing half, roughly
    target = (vis[rng.permutation(len(vis))[:256]] + rng.normal(0, 5e-4, (256, 3))).astype(np.float32)
    init = np.eye(4)
    init[:3, :3] = Rotation.from_rotvec(rng.normal(0, 0.03, 3)).as_matrix() @ R
    init[:3, 3] = t[rec] + rng.normal(0, 0.003, 3)
    return target, init, posed


def test_icp_refine_matches_oracle():
    models = cases.ycb_models()
    classes = [1, 4, 9, 13, 20, 9]
    cases_ = [_icp_case(c, i, 100 + i) for i, c in enumerate(classes)]
    target = np.stack([c[0] for c in cases_])
    init = np.stack([c[1] for c in cases_])
    init[5] = np.eye(4)                                              # nothing within the radius: identity updates
    src6 = np.concatenate([models, np.zeros_like(models)], axis=2)   # [21,2048,6] like obj_models.tfrecords
    T, fit, rmse, iters = EV.icp_refine(cu(src6), cu(target), cu(init), source_of_seg=classes)
    T, fit, rmse, iters = np_(T), np_(fit), np_(rmse), np_(iters)
    for s, c in enumerate(classes):
        oT, ofit, ormse, oit = E.icp_refine(models[c], target[s], init[s])
        np.testing.assert_allclose(T[s], oT, atol=1e-6, err_msg=f"segment {s}")
        assert abs(fit[s] - ofit) < 1e-3 and abs(rmse[s] - ormse) < 1e-6, (s, fit[s], ofit, rmse[s], ormse)
        assert abs(int(iters[s]) - oit) <= 2, (s, iters[s], oit)
        np.testing.assert_allclose(T[s, :3, :3] @ T[s, :3, :3].T, np.eye(3), atol=1e-12)
    assert (T[5] == np.eye(4)).all() and fit[5] == 0.0 and iters[5] == 10
    # property: the refinement reduces the model-to-ground-truth error of a perturbed pose
    for s in range(5):
        m = models[classes[s]].astype(np.float64)
        before = np.abs(m @ init[s][:3, :3].T + init[s][:3, 3] - cases_[s][2]).mean()
        after = np.abs(m @ T[s][:3, :3].T + T[s][:3, 3] - cases_[s][2]).mean()
        assert after < before, (s, before, after)


def test_icp_one_call_equals_registration_icp():
    models = cases.ycb_models()
    target, init, _ = _icp_case(9, 3, 55)
    T, fit, rmse, iters = EV.icp_refine(cu(models[9:10]), cu(target[None]), cu(init[None]), radius=0.015, outer=1,
                                        max_iteration=30)
    oT, ofit, ormse, oit = E.registration_icp(models[9], target, 0.015, init)
    np.testing.assert_allclose(np_(T)[0], oT, atol=1e-7)
    assert abs(float(fit[0]) - ofit) < 1e-9 and abs(float(rmse[0]) - ormse) < 1e-9 and int(iters[0]) == oit


def test_pose_to_matrix_is_rodrigues():
    rng = np.random.default_rng(2)
    a, t = rng.standard_normal((4, 3)).astype(np.float32), rng.standard_normal((4, 3)).astype(np.float32)
    T = np_(EV.pose_to_matrix(cu(a), cu(t)))
    for i in range(4):
        np.testing.assert_allclose(T[i, :3, :3], Rotation.from_rotvec(a[i].astype(np.float64)).as_matrix(), atol=1e-12)
        np.testing.assert_allclose(T[i, :3, 3], t[i].astype(np.float64))


def test_front_end_and_icp_against_committed_golden_vectors():
    """The same golden file the CPU suite pins the oracle with (tests/golden/eval_golden.npz)."""
    import os
    g = np.load(os.path.join(cases.GOLDEN, "eval_golden.npz"))
    depth, label = cases.eval_golden_frame()
    classes = list(cases.EVAL_GOLDEN_CLASSES)
    fe = EV.SegmentFrontEnd(cu(depth[None]), cu(label[None]), cu(E.YCBV_INTRINSICS[None]),
                            cu(np.full(21, 0.2, np.float32)), cap=20000)
    e = fe.extract([0] * len(classes), classes)
    idx, n_in = EV._radius_outliers(e["xyz_org_distance_filtered"], e["num_point_after_filter"], 100, 0.02, 512)
    first = cu(np.full(len(classes), 5, np.int32))
    fps, _ = EV._fps_seeded(e["xyz_org_distance_filtered"], e["num_point_after_filter"], first, 64)
    for s, c in enumerate(classes):
        n_flt, n_inl = int(e["num_point_after_filter"][s]), int(n_in[s])
        num_valid = n_inl - int(idx[s, 0] == 0)
        assert (g[f"seg{c}_counts"] == [int(e["n_org"][s]), n_flt, n_inl, num_valid]).all()
        assert (g[f"seg{c}_mean"] == np_(e["mean"][s])).all()
        assert (g[f"seg{c}_pix_head"] == np_(e["pix"][s, :64])).all()
        assert (g[f"seg{c}_inlier_head"] == np_(idx[s, :64])).all()
        assert (g[f"seg{c}_fps"] == np_(fps[s])).all()
    model, target, init = cases.eval_golden_icp_case()
    T, fit, rmse, it = EV.icp_refine(cu(model[None]), cu(target[None]), cu(init[None]))
    np.testing.assert_allclose(np_(T)[0], g["icp_T"], atol=1e-6)
    assert abs(float(fit[0]) - g["icp_stats"][0]) < 1e-3 and abs(float(rmse[0]) - g["icp_stats"][1]) < 1e-6
    assert abs(int(it[0]) - int(g["icp_stats"][2])) <= 2


def test_add_and_add_s_metrics_match_oracle():
    """ADD / ADD-S for a batch of segments of different classes: ground-truth pose vs a perturbed
    prediction, vs the pose itself (both metrics 0) and vs a 180-degree flip of a near-symmetric object (ADD >> ADD-S)."""
    models = cases.ycb_models()
    t, a, c = cases.ycb_poses()
    per = len(c) // 21
    rng = np.random.default_rng(3)
    classes = [0, 3, 7, 12, 20, 5]
    Tg, Tp = [], []
    for i, cls in enumerate(classes):
        rec = cls * per + i
        G = np.eye(4); G[:3, :3] = Rotation.from_rotvec(a[rec].astype(np.float64)).as_matrix(); G[:3, 3] = t[rec]
        P = np.eye(4)
        if i == 4:
            P = G.copy()                                                   # perfect prediction
        elif i == 5:
            P[:3, :3] = G[:3, :3] @ Rotation.from_rotvec([0, 0, np.pi]).as_matrix(); P[:3, 3] = G[:3, 3]
        else:
            P[:3, :3] = Rotation.from_rotvec(rng.normal(0, 0.05, 3)).as_matrix() @ G[:3, :3]
            P[:3, 3] = G[:3, 3] + rng.normal(0, 0.004, 3)
        Tg.append(G); Tp.append(P)
    src = torch.from_numpy(models).cuda()
    add, adds = EV.add_metrics(src, torch.from_numpy(np.stack(Tg)).cuda(), torch.from_numpy(np.stack(Tp)).cuda(),
                               source_of_seg=classes)
    add, adds = add.cpu().numpy(), adds.cpu().numpy()
    for i, cls in enumerate(classes):
        wa, ws = E.add_metrics(models[cls], Tg[i], Tp[i])
        assert abs(add[i] - wa) <= 1e-6 * max(wa, 1e-3), (i, add[i], wa)
        assert abs(adds[i] - ws) <= 1e-5 * max(ws, 1e-3), (i, adds[i], ws)
        assert adds[i] <= add[i] + 1e-9
    assert add[4] == 0.0 and adds[4] == 0.0
    assert add[5] > 2 * adds[5]
