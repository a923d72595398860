"""Pins oracle/model_ref.py (the float64 restatement every GPU model test is checked against) to THE REFERENCE'S OWN
CODE: /root/reference/models/pointnet_ycb_23_decoder_4.py, utils/tf_util.py and losses/*.py executed in place through
the eager TensorFlow stand-in of oracle/ref_py (TensorFlow 1.12 itself cannot be installed here).

  * where /root/reference exists (the build container): both are run on the same seeded inputs — outputs, losses,
    moving-average updates and the gradient of EVERY trainable variable must agree to float64 rounding;
  * everywhere (the GPU box has no /root/reference): model_ref must reproduce tests/golden/ref_py_golden.npz, which
    tests/golden/make_golden_ref_py.py wrote by running the reference code.
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_ref_py as G  # noqa: E402
from oracle import model_ref as MR, ref_py  # noqa: E402

TOL = 1e-10


def _model_ref(model, p, x, target, trans, axag, train, bn_decay=0.9):
    params = {k: t.clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p.items()}
    upd = {}
    if model == "dgcnn":
        recon, rot, tr, ep = MR.get_model_dgcnn_mean_6d(x, params, train, train, 10, bn_decay, upd)
    else:
        recon, rot, tr, ep = MR.get_model_pn(x, params, train, bn_decay, upd)
    xyz_loss, _ = MR.chamfer_get_loss(recon, target)
    trans_loss, _ = MR.get_translation_error(tr, trans)
    axag_loss, per_rot = MR.get_rotation_error(rot.double(), axag.double())
    total = 1000 * xyz_loss + 10 * trans_loss + axag_loss
    grads = {}
    if train:
        total.backward()
        grads = {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}
    return {"recon": recon, "rot": rot, "trans": tr, "embedding": ep["embedding"], "chamfer": xyz_loss, "trans_loss": trans_loss,
            "rot_loss": axag_loss, "per_rot": per_rot, "total": total}, grads, upd


def _close(a, b, what):
    a = torch.as_tensor(np.asarray(a.detach() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach() if torch.is_tensor(b) else b), dtype=torch.float64)
    err = (a - b).abs().max().item()
    assert err <= TOL * max(1.0, b.abs().max().item()), (what, err)


@pytest.mark.skipif(not ref_py.available(), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("model,b,n,seed,train", G.CASES + [("dgcnn", 5, 48, 21, True), ("pn", 2, 40, 22, False)])
def test_model_ref_equals_the_reference_code_executed_in_place(model, b, n, seed, train):
    tf, model_mod, _, losses = ref_py.load()
    p, x, target, trans, axag = G.case_inputs(model, b, n, seed)
    want, wgrads, wupd = G.run_reference(tf, model_mod, losses, model, p, x, target, trans, axag, train)
    got, ggrads, gupd = _model_ref(model, p, x, target, trans, axag, train)
    for k in want:
        _close(got[k], want[k], k)
    assert sorted(ggrads) == sorted(wgrads)
    for k in wgrads:                                       # every trainable variable
        _close(ggrads[k], wgrads[k], f"grad {k}")
    assert sorted(gupd) == sorted(wupd)
    for k in wupd:
        _close(gupd[k], wupd[k], f"ema {k}")
    # the reference graph created exactly the variables of our parameter set (TF scope names)
    created = set(tf._STATE["created"])
    assert created == {k for k in p if not k.endswith(("ema_mean", "ema_var"))}


@pytest.mark.skipif(not ref_py.available(), reason="/root/reference is not present on this box")
def test_knn_and_edge_feature_helpers_equal_the_reference_tf_util():
    tf, _, tf_util, _ = ref_py.load()
    tf.install({})
    g = torch.Generator().manual_seed(5)
    for shape in ((3, 40, 24), (1, 40, 24)):               # B == 1 takes the re-expand branch (tf_util.py:609-611)
        pc = torch.randn(*shape, generator=g, dtype=torch.float64)
        adj = tf_util.pairwise_xyz_distance(pc)
        _close(MR.pairwise_xyz_distance(pc), adj, "pairwise xyz")
        idx = tf_util.knn(adj, k=10)
        assert torch.equal(MR.knn(adj, 10), idx)
        _close(MR.get_edge_feature(pc, idx), tf_util.get_edge_feature(pc, nn_idx=idx, k=10), "edge feature")
        feat = torch.randn(shape[0], 40, 1, 64, generator=g, dtype=torch.float64)   # layers 2-4: all 64 channels count
        _close(MR.pairwise_xyz_distance(feat), tf_util.pairwise_xyz_distance(feat), "pairwise feature")


def test_model_ref_reproduces_the_committed_reference_vectors():
    z = np.load(os.path.join(HERE, "golden", "ref_py_golden.npz"))
    for model, b, n, seed, train in G.CASES:
        p, x, target, trans, axag = G.case_inputs(model, b, n, seed)
        got, grads, upd = _model_ref(model, p, x, target, trans, axag, train)
        tag = f"{model}_b{b}_{'train' if train else 'eval'}"
        for k, v in got.items():
            _close(v, z[f"{tag}/{k}"], f"{tag}/{k}")
        n_grad = 0
        for key in z.files:
            if key.startswith(f"{tag}/grad_norm/"):
                name = key[len(f"{tag}/grad_norm/"):]
                _close(grads[name].norm(), z[key], key)
                _close(grads[name].reshape(-1)[:16], z[f"{tag}/grad_head/{name}"], key)
                n_grad += 1
            elif key.startswith(f"{tag}/ema/"):
                _close(upd[key[len(f"{tag}/ema/"):]][:8], z[key], key)
        assert n_grad >= (5 if train else 0)
    ax, lab = torch.from_numpy(z["rot_cases/pred"]), torch.from_numpy(z["rot_cases/label"])
    _close(MR.get_rotation_error(ax, lab)[1], z["rot_cases/per"], "rotation corner cases")
    _close(MR.exponential_map(ax), z["rot_cases/expmap"], "exponential map (Taylor branch)")
