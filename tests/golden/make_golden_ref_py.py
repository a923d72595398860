"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN model / loss code (models/pointnet_ycb_23_decoder_4.py,
utils/tf_util.py, losses/*.py under /root/reference, through the eager TensorFlow stand-in of oracle/ref_py) on seeded
inputs.  Run in the build container:  python tests/golden/make_golden_ref_py.py  ->  tests/golden/ref_py_golden.npz
The GPU box has no /root/reference: there the vectors pin oracle/model_ref.py (tests/test_ref_py_pins_model_oracle.py)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import model_ref as MR, ref_py  # noqa: E402


def case_inputs(model, b, n, seed):
    layers = list(MR.DGCNN_LAYERS) if model == "dgcnn" else MR.pn_layers(24, n)
    if model == "dgcnn":
        layers[7] = ("dgcnn_output", 1024, n * 12, False)
    p = {k: t.double() for k, t in MR.init_params(layers, seed=seed, perturb=True).items()}
    g = torch.Generator().manual_seed(seed + 100)
    xyz = torch.randn(b, n, 3, generator=g, dtype=torch.float64) * 0.05
    onehot = torch.nn.functional.one_hot(torch.randint(0, 21, (b,), generator=g), 21).double()
    x = torch.cat([xyz - xyz.mean(1, keepdim=True), onehot[:, None, :].expand(b, n, 21)], 2).contiguous()
    target = torch.randn(b, 4 * n, 3, generator=g, dtype=torch.float64) * 0.05
    trans = torch.randn(b, 3, generator=g, dtype=torch.float64) * 0.1
    axag = torch.randn(b, 3, generator=g, dtype=torch.float64)
    return p, x, target, trans, axag


def run_reference(tf, model_mod, losses, model, p, x, target, trans, axag, train, bn_decay=0.9):
    """The reference's graph section train_cloudAAE_ycbv.py:228-268 with its own functions."""
    params = {k: t.clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p.items()}
    upd = {}
    tf.install(params, ema_updates=upd)
    with ref_py.quiet():
        if model == "dgcnn":
            recon, rot, tr, ep = model_mod.get_model_dgcnn_mean_6d(x, train, train, 10, bn_decay=bn_decay)
        else:
            recon, rot, tr, ep = model_mod.get_model_pn(x, train, bn_decay=bn_decay)
    xyz_loss, _ = losses["chamfer_loss"].get_loss(recon, target)
    trans_loss, _ = losses["trans_distance"].get_translation_error(tr, trans)
    axag_loss, per_rot = losses["angular_distance_taylor"].get_rotation_error(rot.double(), axag.double())
    total = 1000 * xyz_loss + 10 * trans_loss + axag_loss
    grads = {}
    if train:
        total.backward()
        grads = {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}
    return {"recon": recon, "rot": rot, "trans": tr, "embedding": ep["embedding"], "chamfer": xyz_loss, "trans_loss": trans_loss,
            "rot_loss": axag_loss, "per_rot": per_rot, "total": total}, grads, upd


CASES = [("dgcnn", 3, 32, 11, True), ("dgcnn", 1, 32, 12, False), ("dgcnn", 2, 32, 13, False), ("pn", 3, 32, 14, True)]

if __name__ == "__main__":
    tf, model_mod, tf_util, losses = ref_py.load()
    out = {}
    for model, b, n, seed, train in CASES:
        p, x, target, trans, axag = case_inputs(model, b, n, seed)
        res, grads, upd = run_reference(tf, model_mod, losses, model, p, x, target, trans, axag, train)
        tag = f"{model}_b{b}_{'train' if train else 'eval'}"
        for k, v in res.items():
            out[f"{tag}/{k}"] = v.detach().numpy()
        for k in sorted(grads)[::5]:                     # every fifth gradient, as a norm and a 16-element sample
            g = grads[k].detach().reshape(-1)
            out[f"{tag}/grad_norm/{k}"] = np.asarray(g.norm().item())
            out[f"{tag}/grad_head/{k}"] = g[:16].numpy()
        for k, v in upd.items():
            out[f"{tag}/ema/{k}"] = v.detach().numpy()[:8]
    # rotation loss corner cases (Taylor branch, clip)
    rl = losses["angular_distance_taylor"]
    ax = torch.tensor([[1e-3, 2e-3, -1e-3], [0.05, 0.0, 0.0], [3.0, 0.2, -0.1], [0.0, 0.0, 3.14159]], dtype=torch.float64)
    lab = torch.tensor([[1e-3, 2e-3, -1e-3], [0.0, 0.05, 0.0], [-3.0, -0.2, 0.1], [0.0, 0.0, 0.0]], dtype=torch.float64)
    out["rot_cases/pred"], out["rot_cases/label"] = ax.numpy(), lab.numpy()
    out["rot_cases/per"] = rl.get_rotation_error(ax, lab)[1].numpy()
    out["rot_cases/expmap"] = rl.exponential_map(ax).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_py_golden.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "ref_py_golden.npz")), "bytes")
