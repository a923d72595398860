"""Shortlist sizes of the tensor-core kNN screen on the features of a real train step (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cloudaae_b200 import _capi
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
from cloudaae_b200.train import CloudAAETrainer
dev = torch.device("cuda", 0)
B, N = 128, 256
tr = CloudAAETrainer(batch_size=B, num_point=N, device=dev, seed=0)
syn = SegmentSynthesizer(load_models_xyz(device=dev), B, N, seed=1234)
lib, st = _capi.lib(), torch.cuda.current_stream().cuda_stream
eng = tr.engine
for bi, bt in enumerate(bench.pose_batches(B, seed=0)[:4]):
  c, ax, tl = (torch.from_numpy(bt[k]).to(dev) for k in bench.TRAIN_KEYS)
  tr.train_step_online(syn, c, ax, tl); torch.cuda.synchronize()
  print("batch", bi, "num_vis min", int(syn.num_vis.min()), "flags", int(eng.knn_flags.sum()))
  for name, feat, ld, ch in (("L1 xyz", tr.x, eng.D, 3), ("L2", eng.hcat, 320, 64), ("L3", eng.hcat[:, 64:], 320, 64), ("L4", eng.hcat[:, 128:], 320, 64)):
      idx = torch.empty(B * N, 10, dtype=torch.int32, device=dev); cnt = torch.zeros(B * N, dtype=torch.int32, device=dev)
      _capi.check(lib.caae_debug_knn_shortlist(B, N, ch, 10, feat.data_ptr(), ld, idx.data_ptr(), cnt.data_ptr(), st), "dbg")
      torch.cuda.synchronize()
      cf = cnt.float()
      percloud = cnt.view(B, N).float().max(1).values
      f = feat[:, :ch].reshape(B, N, ch) if feat.dim() == 2 else feat[:, :, :ch]
      fc = f - f.mean(1, keepdim=True)
      print(f"{name}: shortlist mean {cf.mean():.1f} median {cf.median():.0f} p99 {cf.quantile(0.99):.0f} max {cf.max():.0f}; rows > 32: {(cnt > 32).float().mean():.4f}; "
            f"rows > 48: {(cnt > 48).sum().item()}; clouds with a row > 32: {(percloud > 32).sum().item()}/{B}, > 48: {(percloud > 48).sum().item()}; top cloud maxima {sorted(percloud.tolist())[-4:]}")
