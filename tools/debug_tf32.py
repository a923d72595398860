import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cloudaae_b200 import _capi
lib = _capi.lib()
def run(ta, tb, M, N, K, A, B):
    C = torch.zeros(M, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.caae_gemm_tf32(ta, tb, M, N, K, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], C.data_ptr(), N, None, 0, st)
    torch.cuda.synchronize()
    return rc, C
torch.manual_seed(0)
for (M, N, K) in [(128, 128, 32), (128, 128, 8), (128, 128, 64), (256, 256, 128)]:
    for ta in (0, 1):
        for tb in (0, 1):
            A = torch.randn((K, M) if ta else (M, K), device="cuda")
            B = torch.randn((N, K) if tb else (K, N), device="cuda")
            rc, C = run(ta, tb, M, N, K, A, B)
            want = (A.T if ta else A).double() @ (B.T if tb else B).double()
            err = (C.double() - want).abs().max().item()
            print(f"M{M} N{N} K{K} ta={ta} tb={tb} rc={rc} maxerr={err:.4f} ref_max={want.abs().max().item():.2f}")
# structured probe for the failing layout: A = one-hot rows to see which B elements are picked up
M, N, K = 128, 128, 32
for ta, tb in [(0, 0), (1, 1)]:
    A = torch.zeros((K, M) if ta else (M, K), device="cuda")
    B = torch.zeros((N, K) if tb else (K, N), device="cuda")
    # B[k, n] = k*1000 + n ; A selects k = 5 for row 0, k = 17 for row 1
    kk = torch.arange(K, device="cuda").float()[:, None]; nn = torch.arange(N, device="cuda").float()[None, :]
    Bl = kk * 1000 + nn
    B.copy_(Bl.T if tb else Bl)
    Al = torch.zeros(M, K, device="cuda"); Al[0, 5] = 1; Al[1, 17] = 1; Al[2, 0] = 1
    A.copy_(Al.T if ta else Al)
    rc, C = run(ta, tb, M, N, K, A, B)
    print("probe ta,tb", ta, tb, "row0[:8]", C[0, :8].tolist(), "row0[32:36]", C[0, 32:36].tolist())
    print("   row1[:4]", C[1, :4].tolist(), "row2[:4]", C[2, :4].tolist())
