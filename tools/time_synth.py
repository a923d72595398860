"""Times the on-line synthesis stage (B=128) with CUDA events; prints per-kernel split. Debug aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
from cloudaae_b200 import _capi
B = 128
dev = torch.device("cuda", 0)
syn = SegmentSynthesizer(load_models_xyz(device=dev), B, 256, seed=1234)
pool = bench.pose_batches(B, seed=0, pool=4)
pool_d = [{k: torch.from_numpy(v).to(dev) for k, v in bt.items()} for bt in pool]
def run(i):
    bt = pool_d[i % 4]
    syn.synthesize(*[bt[k] for k in bench.TRAIN_KEYS])
for i in range(3): run(i)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(20): run(i)
e.record(); torch.cuda.synchronize()
print("synthesis ms/batch: %.3f" % (s.elapsed_time(e) / 20))
p = _capi.ptr
n = syn.nm + syn.no
st = torch.cuda.current_stream().cuda_stream
lib = _capi.lib()
for name, args in (("hpr occluded", (B, n, p(syn.flip_all), p(syn.points), n, syn.N, p(syn.pad_u), p(syn.visible), p(syn.num_vis), None)),
                   ("hpr org", (B, syn.nm, p(syn.flip_org), p(syn.points), n, 4 * syn.N, p(syn.pad_u_org), p(syn.target), p(syn.num_vis_org), None))):
    s.record()
    for i in range(20): lib.caae_hpr_select(*args, st)
    e.record(); torch.cuda.synchronize()
    print("%s ms: %.3f" % (name, s.elapsed_time(e) / 20))
print("num_vis mean", syn.num_vis.float().mean().item(), "org", syn.num_vis_org.float().mean().item())
s.record()
for i in range(20):
    lib.caae_hpr_select_pair(B, n, p(syn.flip_all), syn.N, p(syn.pad_u), p(syn.visible), p(syn.num_vis), syn.nm, p(syn.flip_org),
                             4 * syn.N, p(syn.pad_u_org), p(syn.target), p(syn.num_vis_org), p(syn.points), n, st)
e.record(); torch.cuda.synchronize()
print("hpr pair ms: %.3f" % (s.elapsed_time(e) / 20))
