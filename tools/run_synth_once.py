import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ycb_poses.npz"))
B = 128
sel = np.random.default_rng(0).integers(0, len(z["class_id"]), B)
syn = SegmentSynthesizer(load_models_xyz(), B, 256, seed=1)
c = torch.from_numpy(z["class_id"][sel].astype(np.int32)).cuda(); a = torch.from_numpy(z["axisangle"][sel]).cuda(); t = torch.from_numpy(z["translation"][sel]).cuda()
for _ in range(3):
    syn.synthesize(c, a, t)
torch.cuda.synchronize()
