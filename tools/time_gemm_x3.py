"""Timing of the split-precision forward GEMMs next to the single-pass TF32 ones (warm, back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cloudaae_b200 import _capi
lib = _capi.lib()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for (M, N, K, lda) in ((32768, 1024, 320, 320), (32768, 128, 64, 320), (32768, 256, 64, 320), (32768, 1024, 128, 128)):
    A = torch.randn(M, lda, device="cuda"); B = torch.randn(K, N, device="cuda"); C = torch.empty(M, N, device="cuda")
    A_lo = torch.empty_like(A); B_lo = torch.empty_like(B)
    lib.caae_split_tf32(M, lda, A.data_ptr(), lda, A_lo.data_ptr(), lda, st)
    lib.caae_split_tf32(K, N, B.data_ptr(), N, B_lo.data_ptr(), N, st)
    parts = torch.zeros(4 * 128 * 2 * N, dtype=torch.float64, device="cuda")
    t1 = timeit(lambda: lib.caae_gemm_tf32(0, 0, M, N, K, A.data_ptr(), lda, B.data_ptr(), N, C.data_ptr(), N, None, 0, st))
    t3 = timeit(lambda: lib.caae_gemm_tf32x3(0, 0, M, N, K, A.data_ptr(), A_lo.data_ptr(), lda, B.data_ptr(), B_lo.data_ptr(), N,
                                             C.data_ptr(), N, None, 0, None, st))
    ts = timeit(lambda: lib.caae_split_tf32(M, K, A.data_ptr(), lda, A_lo.data_ptr(), lda, st))
    print(f"M={M} N={N} K={K}: tf32 {t1:.1f} us, tf32x3 {t3:.1f} us ({2.0*M*N*K*3/t3/1e6:.0f} TFLOP/s of tensor work), split(A) {ts:.1f} us")
