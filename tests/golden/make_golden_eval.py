"""Generate golden vectors for the evaluation front end and the ICP refinement (run in the build container).

    python tests/golden/make_golden_eval.py

Outputs tests/golden/eval_golden.npz.  Inputs are regenerated from seeds by the tests (tests/cases.py:
eval_golden_frame, eval_golden_icp_case); only expected OUTPUTS of oracle/evaluation.py are stored, per class c of
the frame: seg{c}_counts = [n_org, n_filt, n_inlier, num_valid], seg{c}_mean, seg{c}_pix_head (first 64 pixel ids
of the distance-filtered list), seg{c}_inlier_head (first 64 inlier ids), seg{c}_fps (FPS_random, K = 64, first
index 5, on the distance-filtered cloud); icp_T / icp_stats = [fitness, inlier_rmse, iterations] of the 10-round
refinement.  The reference holds no vector for this path (parity unpinned, see oracle/evaluation.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import evaluation as E  # noqa: E402
import cases  # noqa: E402


def main():
    out = {}
    depth, label = cases.eval_golden_frame()
    out["frame_checksum"] = np.array([int(depth.astype(np.int64).sum()), int(label.astype(np.int64).sum())])
    for c in cases.EVAL_GOLDEN_CLASSES:
        org, flt, pix, mean = E.segment_extract(depth, label, E.YCBV_INTRINSICS, c, 0.2)
        idx = E.get_outlier_idx(flt)
        out[f"seg{c}_counts"] = np.array([len(org), len(flt), len(idx), E.num_valid_points(idx)])
        out[f"seg{c}_mean"] = mean
        out[f"seg{c}_pix_head"] = pix[:64]
        out[f"seg{c}_inlier_head"] = idx[:64]
        out[f"seg{c}_fps"] = E.FPS_random(flt, 64, 5)
    model, target, init = cases.eval_golden_icp_case()
    T, fit, rmse, it = E.icp_refine(model, target, init)
    out["icp_T"] = T
    out["icp_stats"] = np.array([fit, rmse, it])
    np.savez_compressed(os.path.join(HERE, "eval_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
