import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cloudaae_b200 import _capi
lib = _capi.lib(); st = torch.cuda.current_stream().cuda_stream
b, n, c, ld = 16, 256, 64, 320
torch.manual_seed(0)
x = torch.relu(torch.randn(b, n, ld, device="cuda") * 0.01 + 1.0)
for i, V in enumerate((3, 30, 60, 100)):
    pick = torch.randint(0, V, (n - V,), device="cuda")
    x[i, V:] = x[i, pick]
idx = torch.empty(b, n, 10, dtype=torch.int32, device="cuda")
for _ in range(3):
    lib.caae_knn(b, n, c, 10, x.data_ptr(), ld, idx.data_ptr(), st)
torch.cuda.synchronize()
