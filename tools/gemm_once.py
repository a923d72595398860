"""The three dgcnn_agg contractions once each (after a warm-up) — target of single-kernel ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cloudaae_b200 import _capi
lib = _capi.lib(); st = torch.cuda.current_stream().cuda_stream
R = 32768
X = torch.randn(R, 320, device="cuda"); W = torch.randn(320, 1024, device="cuda") * 0.05
Y = torch.empty(R, 1024, device="cuda"); dX = torch.empty(R, 320, device="cuda"); dW = torch.empty(320, 1024, device="cuda")
def g(ta, tb, M, N, K, A, lda, B, ldb, C, ldc):
    _capi.check(lib.caae_gemm_tf32(ta, tb, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, C.data_ptr(), ldc, None, 0, st), "gemm")
for _ in range(2):
    g(0, 0, R, 1024, 320, X, 320, W, 1024, Y, 1024)
    g(0, 1, R, 320, 1024, Y, 1024, W, 1024, dX, 320)
    g(1, 0, 320, 1024, R, X, 320, Y, 1024, dW, 1024)
torch.cuda.synchronize()
