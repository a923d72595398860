"""Collect per-cloud HPR cost (cycles) with the class id, for both problems, over several batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cloudaae_b200 import _capi
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
B = 128
dev = torch.device("cuda", 0)
syn = SegmentSynthesizer(load_models_xyz(device=dev), B, 256, seed=1234)
lib = _capi.lib(); p = _capi.ptr; n = syn.nm + syn.no; st = torch.cuda.current_stream().cuda_stream
rows = []
for bi, bt_h in enumerate(bench.pose_batches(B, seed=0, pool=8)):
    bt = {k: torch.from_numpy(v).to(dev) for k, v in bt_h.items()}
    syn.synthesize(*[bt[k] for k in bench.TRAIN_KEYS])
    torch.cuda.synchronize()
    for prob, args in ((0, (B, n, p(syn.flip_all), p(syn.points), n, syn.N, p(syn.pad_u), p(syn.visible), p(syn.num_vis), None)),
                       (1, (B, syn.nm, p(syn.flip_org), p(syn.points), n, 4 * syn.N, p(syn.pad_u_org), p(syn.target), p(syn.num_vis_org), None))):
        lib.caae_hpr_select(*args, st); torch.cuda.synchronize()
        buf = np.zeros((512, 8), np.int64)
        assert lib.caae_debug_hpr_timing(buf.ctypes.data) == 0
        tot = buf[:B, :5].sum(1)
        nv = (syn.num_vis if prob == 0 else syn.num_vis_org).cpu().numpy()
        for i in range(B):
            rows.append((bi, prob, int(bt_h["class_id"][i]), int(tot[i]), int(nv[i]), *[int(x) for x in buf[i]], float(bt_h["translation"][i][2])))
np.save("gpurun_out/hpr_cost.npy", np.array(rows, np.float64))
a = np.array(rows, np.float64)
for prob in (0, 1):
    s = a[a[:, 1] == prob]
    print("problem", prob, "mean %.0f max %.0f std %.0f" % (s[:, 3].mean(), s[:, 3].max(), s[:, 3].std()))
    resid = s[:, 3].copy()
    for c in range(21):
        m = s[:, 2] == c
        if m.any():
            resid[m] -= s[m, 3].mean()
    print("   residual std after class mean: %.0f" % resid.std())
