import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cloudaae_b200 import _capi
lib = _capi.lib(); st = torch.cuda.current_stream().cuda_stream
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for (b, n, c, ld, spread) in ((128, 256, 64, 320, 0.3), (128, 256, 64, 320, 0.01), (128, 256, 3, 24, 0.3), (8, 256, 64, 320, 0.3)):
    x = torch.relu(torch.randn(b, n, ld, device="cuda") * spread + 1.0)   # common mean 1, per-point spread
    if spread == 0.01 and c == 64:   # a few heavily padded clouds (V visible points + random repeats), as the synthesis produces
        for i, V in enumerate((3, 30, 60, 100)):
            pick = torch.randint(0, V, (n - V,), device="cuda")
            x[i, V:] = x[i, pick]
    idx = torch.empty(b, n, 10, dtype=torch.int32, device="cuda")
    t_tc = timeit(lambda: lib.caae_knn(b, n, c, 10, x.data_ptr(), ld, idx.data_ptr(), st))
    t_ff = timeit(lambda: lib.caae_knn_ffma(b, n, c, 10, x.data_ptr(), ld, idx.data_ptr(), st))
    print(f"b={b} n={n} c={c} spread={spread}: caae_knn (tensor-core screen, routed) {t_tc:.1f} us, FFMA kernel {t_ff:.1f} us")
