"""Pins oracle/synthesis.py (occluder, spherical flips, convex-hull visibility) to THE REFERENCE'S OWN utilities
(utils/generate_occluder.py, utils/hidden_point_removal.py) executed in place through oracle/ref_py — where
/root/reference exists — and to the vectors that run left under tests/golden/ everywhere else."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_ref_py_synth as G  # noqa: E402
from oracle import ref_py, synthesis as S  # noqa: E402


def _oracle(model, ax, tr, zc, zp):
    P = S.transform_object_model(model[None], ax[None], tr[None])
    occ = S.spherical_occluder(tr[None, 2], zc, zp)
    fl, org = S.spherical_flip(np.concatenate([P, occ], 1))
    vis, num, _ = S.convex_hull_visible(fl, org)
    fl2, org2 = S.spherical_flip(P)
    vis2, num2, _ = S.convex_hull_visible(fl2, org2)
    return occ[0], fl[0], fl2[0], vis[0, :num[0]], vis2[0, :num2[0]], int(num[0]), int(num2[0])


def _check(got, occ, fl_head, fl2_head, vis, vis2, n, n2):
    o_occ, o_fl, o_fl2, o_vis, o_vis2, o_n, o_n2 = got
    assert np.array_equal(o_occ, occ)                                # occluder: bit-exact
    for a, b in ((o_fl[:len(fl_head)], fl_head), (o_fl2[:len(fl2_head)], fl2_head)):
        # flips: a few fp32 ulps — the association order inside tf.norm's 3-element reduction (torch here, Eigen in
        # TensorFlow, NumPy in the oracle) moves |p| by 1 ulp and 2 (R - |p|) p / |p| + p carries it through four roundings
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
    assert (o_n, o_n2) == (n, n2)                                    # Qhull on both: identical visible sets
    assert np.array_equal(o_vis, vis) and np.array_equal(o_vis2, vis2)


@pytest.mark.skipif(not ref_py.available(), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("rec,seed", G.SAMPLES + [(5, 9), (3000, 4)])
def test_synthesis_oracle_equals_the_reference_utilities_executed_in_place(rec, seed):
    tf, H, GO = G.load_utils()
    inp = G.sample_inputs(rec, seed)
    r = G.run_reference(tf, H, GO, *inp)
    n, n2 = int(r["num_vis_point"][0]), int(r["num_vis_point_org"][0])
    _check(_oracle(*inp), r["occluder"][0], r["flippedPoints"][0], r["flippedPoints_org"][0], r["visiblePoints"][0, :n],
           r["visiblePoints_org"][0, :n2], n, n2)
    # the reference pads with random repeats of visible points only, up to P + 1 rows
    pad = r["visiblePoints"][0, n:]
    assert len(r["visiblePoints"][0]) == 2449 and all((pad[i] == r["visiblePoints"][0, :n]).all(1).any() for i in range(0, len(pad), 97))


def test_synthesis_oracle_reproduces_the_committed_reference_vectors():
    z = np.load(os.path.join(HERE, "golden", "ref_py_synth_golden.npz"))
    for rec, seed in G.SAMPLES:
        t = f"rec{rec}"
        n, n2 = (int(v) for v in z[f"{t}/num_vis"])
        _check(_oracle(*G.sample_inputs(rec, seed)), z[f"{t}/occluder"], z[f"{t}/flipped_head"], z[f"{t}/flipped_org_head"],
               z[f"{t}/visible"], z[f"{t}/visible_org"], n, n2)
