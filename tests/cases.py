"""Seeded inputs shared by the parity tests and the golden-vector generators."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ycb_models() -> np.ndarray:
    return np.load(os.path.join(GOLDEN, "ycb_models_xyz.npy"))


def ycb_poses():
    z = np.load(os.path.join(GOLDEN, "ycb_poses.npz"))
    return z["translation"], z["axisangle"], z["class_id"]


def nnd_smoke_inputs(nclouds: int = 2):
    """The reference smoke test's arrays (tf_nndistance.py:42-49), first `nclouds` clouds.
    The legacy RandomState stream is frozen across NumPy versions."""
    rs = np.random.RandomState(100)
    xyz1 = rs.randn(32, 16384, 3).astype("float32")
    xyz2 = rs.randn(32, 1024, 3).astype("float32")
    return np.ascontiguousarray(xyz1[:nclouds]), np.ascontiguousarray(xyz2[:nclouds])


def nnd_smoke_grads(nclouds: int = 2):
    rng = np.random.default_rng(7)
    return (rng.standard_normal((nclouds, 16384)).astype(np.float32),
            rng.standard_normal((nclouds, 1024)).astype(np.float32))


def posed_ycb_clouds(record: int = 0) -> np.ndarray:
    """The 21 YCB models, each posed by pose record `record` of its own class -> f32[21,2048,3]."""
    from oracle.synthesis import transform_object_model
    t, a, c = ycb_poses()
    per = len(c) // 21
    sel = np.arange(21) * per + record
    assert (c[sel] == np.arange(21)).all()
    return transform_object_model(ycb_models(), a[sel], t[sel])


def fps_ycb_inputs() -> np.ndarray:
    return posed_ycb_clouds(0)


def fps_ties_inputs() -> np.ndarray:
    """2048 points whose second half duplicates the first: every distance tie is between k and
    k+1024, which the reference resolves by (k mod 512) first — not by lowest index."""
    rng = np.random.default_rng(11)
    half = rng.uniform(-0.1, 0.1, (3, 1024, 3)).astype(np.float32)
    # make ties land across different (k mod 512) classes too: shuffle the duplicate half
    perm = np.random.default_rng(12).permutation(1024)
    return np.concatenate([half, half[:, perm]], axis=1)


def random_clouds(seed: int, b: int, n: int, scale: float = 0.1) -> np.ndarray:
    return (np.random.default_rng(seed).standard_normal((b, n, 3)) * scale).astype(np.float32)


# ---- evaluation front end / ICP (oracle/evaluation.py, tests/golden/eval_golden.npz) ---------------------
EVAL_GOLDEN_CLASSES = (0, 3, 7)


def eval_golden_frame():
    """One synthetic 480 x 640 frame: YCB models 0, 3, 7 posed by record 0 of their class (depth u16, label u8)."""
    from cloudaae_b200.data.synthetic_frames import render_frame
    clouds = posed_ycb_clouds(0)
    return render_frame(clouds[list(EVAL_GOLDEN_CLASSES)], list(EVAL_GOLDEN_CLASSES), splat=1, seed=0)


def eval_golden_icp_case(cls: int = 9, seed: int = 55):
    """A 256-point segment cut from the camera-facing half of a posed model + a perturbed initial pose."""
    from scipy.spatial.transform import Rotation
    models = ycb_models()
    t, a, c = ycb_poses()
    rec = cls * (len(c) // 21) + 3
    rng = np.random.default_rng(seed)
    R = Rotation.from_rotvec(a[rec].astype(np.float64)).as_matrix()
    posed = models[cls].astype(np.float64) @ R.T + t[rec]
    vis = posed[posed[:, 2] < np.median(posed[:, 2])]
    target = (vis[rng.permutation(len(vis))[:256]] + rng.normal(0, 5e-4, (256, 3))).astype(np.float32)
    init = np.eye(4)
    init[:3, :3] = Rotation.from_rotvec(rng.normal(0, 0.03, 3)).as_matrix() @ R
    init[:3, 3] = t[rec] + rng.normal(0, 0.003, 3)
    return models[cls], target, init
