"""Does tcgen05.mma kind::tf32 truncate or round its fp32-container operands?  (decides how the low part of a
split-precision operand must be formed)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cloudaae_b200 import _capi
lib = _capi.lib()
M = N = 128; K = 32
st = torch.cuda.current_stream().cuda_stream
for name, val in (("1+3*2^-12", 1 + 3 * 2.0 ** -12), ("1+2^-11 (tie)", 1 + 2.0 ** -11), ("1+2^-11+2^-23", 1 + 2.0 ** -11 + 2.0 ** -23),
                  ("-(1+3*2^-12)", -(1 + 3 * 2.0 ** -12)), ("1+2^-10+2^-11 (tie, odd)", 1 + 2.0 ** -10 + 2.0 ** -11)):
    for which in ("A", "B"):
        A = torch.full((M, K), val if which == "A" else 1.0, device="cuda")
        B = torch.full((K, N), val if which == "B" else 1.0, device="cuda")
        C = torch.zeros(M, N, device="cuda")
        _capi.check(lib.caae_gemm_tf32(0, 0, M, N, K, A.data_ptr(), K, B.data_ptr(), N, C.data_ptr(), N, None, 0, st), "g")
        torch.cuda.synchronize()
        print(f"{name:28s} in {which}: C/K = {C[0,0].item()/K!r}   (x = {val!r})")
