"""caae_knn_part(1) on the layer-2 features of a real train step (for ncu)."""
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
bt = bench.pose_batches(B, seed=0)[0]
c, ax, tl = (torch.from_numpy(bt[k]).to(dev) for k in bench.TRAIN_KEYS)
tr.train_step_online(syn, c, ax, tl); torch.cuda.synchronize()
lib, st = _capi.lib(), torch.cuda.current_stream().cuda_stream
eng = tr.engine
idx = torch.empty(B * N, 10, dtype=torch.int32, device=dev)
for _ in range(3):
    lib.caae_knn_part(1, eng.knn_flags.data_ptr(), B, N, 64, 10, eng.hcat.data_ptr(), 320, idx.data_ptr(), st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    lib.caae_knn_part(1, eng.knn_flags.data_ptr(), B, N, 64, 10, eng.hcat.data_ptr(), 320, idx.data_ptr(), st)
b.record(); torch.cuda.synchronize()
print("L2 knn on real features:", a.elapsed_time(b) / 20 * 1e3, "us")
