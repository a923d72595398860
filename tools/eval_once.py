"""Evaluation front end + ICP once (after a warm-up) on synthetic frames — target of ncu captures."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
from cloudaae_b200 import evaluate_cloudAAE_ycbv as EV
from cloudaae_b200.data import synthetic_frames as SF
dev = torch.device("cuda", 0)
clouds = cases.posed_ycb_clouds(0)
fr = [SF.render_frame(clouds[[3 * f, 3 * f + 1, 3 * f + 2]], [3 * f, 3 * f + 1, 3 * f + 2], splat=2, seed=f) for f in range(4)]
fe = EV.SegmentFrontEnd(torch.from_numpy(np.stack([d for d, _ in fr])).to(dev), torch.from_numpy(np.stack([l for _, l in fr])).to(dev),
                        torch.from_numpy(np.tile(SF.YCBV_INTRINSICS, (4, 1))).to(dev), torch.full((21,), 0.2, device=dev), cap=49152)
fos = [f for f in range(4) for _ in range(3)]; cos = list(range(12))
for _ in range(2):
    out = fe.run(fos, cos, 256, rng=random.Random(0))
models = torch.from_numpy(cases.ycb_models()).to(dev)
t, a, c = cases.ycb_poses(); per = len(c) // 21
sel = np.arange(12) * per
T0 = EV.pose_to_matrix(torch.from_numpy(a[sel]).to(dev), torch.from_numpy(t[sel]).to(dev)); T0[:, :3, 3] += 0.003
for _ in range(2):
    T, fit, rmse, it = EV.icp_refine(models, out["xyz_inlier"], T0, source_of_seg=cos)
torch.cuda.synchronize()
print("points", out["num_point_after_filter"].tolist(), "icp iters", it.tolist(), "fitness", [round(x, 3) for x in fit.tolist()])
