"""Dense TF32 tensor-core peak of this GPU, measured the way MEASURED_PEAKS.json measures bf16:
torch.matmul 8192^3 with allow_tf32 — best of 10 (burst) and back to back for ~2 s (sustained)."""
import json
import sys
import time

import torch


def measure(seconds: float = 2.0, n: int = 8192):
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    c = torch.empty(n, n, device="cuda")
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a, b, out=c); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    flops = 2.0 * n ** 3
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); iters = 0
    s.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            torch.matmul(a, b, out=c)
        iters += 10
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    sustained = flops * iters / (s.elapsed_time(e) * 1e-3) / 1e12
    return {"tf32_tflops": flops / (best * 1e-3) / 1e12, "tf32_tflops_sustained": sustained,
            "how": f"torch.matmul fp32 {n}^3 allow_tf32: best of 10 (burst), back to back {seconds:.0f} s (sustained)"}


if __name__ == "__main__":
    print(json.dumps(measure(float(sys.argv[1]) if len(sys.argv) > 1 else 2.0)))
