"""Per-phase clock split of the HPR kernel (B=128, both variants). Debug aid, run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cloudaae_b200 import _capi
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
B = 128
dev = torch.device("cuda", 0)
syn = SegmentSynthesizer(load_models_xyz(device=dev), B, 256, seed=1234)
bt = {k: torch.from_numpy(v).to(dev) for k, v in bench.pose_batches(B, seed=0, pool=1)[0].items()}
syn.synthesize(*[bt[k] for k in bench.TRAIN_KEYS])
torch.cuda.synchronize()
lib = _capi.lib(); p = _capi.ptr; n = syn.nm + syn.no; st = torch.cuda.current_stream().cuda_stream
for name, args in (("occluded", (B, n, p(syn.flip_all), p(syn.points), n, syn.N, p(syn.pad_u), p(syn.visible), p(syn.num_vis), None)),
                   ("org", (B, syn.nm, p(syn.flip_org), p(syn.points), n, 4 * syn.N, p(syn.pad_u_org), p(syn.target), p(syn.num_vis_org), None))):
    lib.caae_hpr_select(*args, st); torch.cuda.synchronize()
    buf = np.zeros((512, 8), np.int64)
    assert lib.caae_debug_hpr_timing(buf.ctypes.data) == 0
    t = buf[:B]
    names = ["setup", "phase1", "verify0", "rounds", "select", "nsurv", "nrounds", "resolved0"]
    print(name, "mean:", " ".join(f"{k}={t[:, i].mean():.0f}" for i, k in enumerate(names)))
    print(name, "max :", " ".join(f"{k}={t[:, i].max():.0f}" for i, k in enumerate(names)))
    tot = t[:, :5].sum(1)
    print(name, "total cycles mean %.0f max %.0f (%.3f ms at 1.9 GHz)" % (tot.mean(), tot.max(), tot.max() / 1.9e6))
