"""Golden vectors of the on-line synthesis produced by EXECUTING THE REFERENCE'S OWN utilities
(utils/generate_occluder.py, utils/hidden_point_removal.py — incl. its scipy ConvexHull py_func — under /root/reference,
through the eager TensorFlow stand-in of oracle/ref_py) with scripted normal draws.
    python tests/golden/make_golden_ref_py_synth.py  ->  tests/golden/ref_py_synth_golden.npz"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_py, synthesis as S  # noqa: E402

SAMPLES = [(1234, 0), (77, 1), (5011, 2)]      # (pose record, seed of the occluder draws)


def load_utils():
    tf, *_ = ref_py.load()
    saved_path, saved_tf = list(sys.path), sys.modules.get("tensorflow")
    sys.path[:0] = [ref_py._SHIM, os.path.join(ref_py.REF_ROOT, "utils")]
    sys.modules["tensorflow"] = tf
    try:
        for n in ("hidden_point_removal", "generate_occluder", "sample_pose_in_frustum"):
            sys.modules.pop(n, None)
        H = importlib.import_module("hidden_point_removal")
        GO = importlib.import_module("generate_occluder")
    finally:
        sys.path[:] = saved_path
        if saved_tf is not None:
            sys.modules["tensorflow"] = saved_tf
        else:
            sys.modules.pop("tensorflow", None)
    return tf, H, GO


def sample_inputs(rec, seed):
    models = np.load(os.path.join(HERE, "ycb_models_xyz.npy"))
    z = np.load(os.path.join(HERE, "ycb_poses.npz"))
    rng = np.random.default_rng(seed)
    cls, ax, tr = int(z["class_id"][rec]), z["axisangle"][rec], z["translation"][rec]
    zc = rng.standard_normal((1, 2, 3)).astype(np.float32)
    zp = rng.standard_normal((1, 2, 200, 3)).astype(np.float32)
    return models[cls], ax, tr, zc, zp


def run_reference(tf, H, GO, model, ax, tr, zc, zp):
    """train_cloudAAE_ycbv.py:100-111 with the reference's functions; the posed model comes from the oracle's
    transform (its float64 exponential map is pinned by test_ref_py_pins_model_oracle.py)."""
    tf.install({}, dtype=torch.float32)
    P = S.transform_object_model(model[None], ax[None], tr[None])
    draws = [torch.from_numpy(zc[0, :, d:d + 1].copy()) for d in range(3)]
    draws += [torch.from_numpy(zp[0, blob, :, d:d + 1].copy()) for blob in range(2) for d in range(3)]
    tf.script_normal_draws(draws)
    x = GO.get_random_spherical_occluder({"translation": torch.from_numpy(tr[None].copy())}, "ycbv")
    x["model_xyz_rot_trans"] = torch.from_numpy(P)
    param = torch.tensor([[0.8 * np.pi]], dtype=torch.float32)
    x = H.sphericalFlip(x, torch.zeros(1, 3), param)
    x = H.sphericalFlip_org(x, torch.zeros(1, 3), param)
    np.random.seed(0)
    x = H.hidden_point_removal(x)
    x = H.hidden_point_removal_org(x)
    return {k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in x.items()}


if __name__ == "__main__":
    tf, H, GO = load_utils()
    out = {}
    for rec, seed in SAMPLES:
        r = run_reference(tf, H, GO, *sample_inputs(rec, seed))
        n, no = int(r["num_vis_point"][0]), int(r["num_vis_point_org"][0])
        tag = f"rec{rec}"
        out[f"{tag}/occluder"] = r["occluder"][0]
        out[f"{tag}/flipped_head"] = r["flippedPoints"][0, :64]
        out[f"{tag}/flipped_org_head"] = r["flippedPoints_org"][0, :64]
        out[f"{tag}/num_vis"] = np.asarray([n, no])
        out[f"{tag}/visible"] = r["visiblePoints"][0, :n]             # the deterministic part (before the random padding)
        out[f"{tag}/visible_org"] = r["visiblePoints_org"][0, :no]
    np.savez_compressed(os.path.join(HERE, "ref_py_synth_golden.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "ref_py_synth_golden.npz")), "bytes")
