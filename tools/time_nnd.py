"""nn_distance forward at the step's shape (b = 128, n = m = 1024): us per launch for the queries-per-thread variant
selected by CAAE_NND_Q (unset: the launcher's own choice)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cloudaae_b200 import nn_distance

b, n = 128, 1024
g = torch.Generator("cuda").manual_seed(0)
x = torch.randn(b, n, 3, device="cuda", generator=g)
y = torch.randn(b, n, 3, device="cuda", generator=g)
for _ in range(20):
    nn_distance(x, y)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    s.record()
    for _ in range(50):
        nn_distance(x, y)
    e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e) / 50 * 1e3)
print(f"CAAE_NND_Q={os.environ.get('CAAE_NND_Q', 'auto')}: nn_distance fwd b={b} n=m={n}: {best:.1f} us per call (incl. output allocation)")
