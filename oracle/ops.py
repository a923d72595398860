"""ctypes bindings for the C restatement (oracle/caae_oracle.c) and for the reference's own
compiled code (oracle/_ref/libcloudaae_ref.so, built by oracle/build_ref.sh).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "_build", "libcaae_oracle.so")
_REF_SO = os.path.join(HERE, "_ref", "libcloudaae_ref.so")

_c_f = ctypes.POINTER(ctypes.c_float)
_c_i = ctypes.POINTER(ctypes.c_int)
_int = ctypes.c_int


def build(ref: bool = True) -> None:
    """Compile the restatement (and the reference, when /root/reference exists)."""
    subprocess.check_call(["make", "-C", HERE, "_build/libcaae_oracle.so"] + (["ref"] if ref else []),
                          stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "caae_oracle.c")
        if not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
            build(ref=False)
        _lib = ctypes.CDLL(_ORACLE_SO)
        _lib.oracle_max_threads.restype = _int
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref() -> ctypes.CDLL:
    """The reference's own compiled code; raises if it was never built."""
    global _ref
    if _ref is None:
        if not have_ref():
            build(ref=True)
        if not have_ref():
            raise FileNotFoundError(f"{_REF_SO} missing and /root/reference not available to build it")
        _ref = ctypes.CDLL(_REF_SO)
        _ref.ref_last_error.restype = ctypes.c_char_p
    return _ref


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pf(a):
    return a.ctypes.data_as(_c_f)


def _pi(a):
    return a.ctypes.data_as(_c_i)


# ------------------------------------------------------------------ restatement
def fps(inp, npoint: int, threads: int = 1) -> np.ndarray:
    """farthest_point_sample(npoint, inp) — tf_sampling_g.cu:105-170. inp f32[b,n,3] -> i32[b,npoint]."""
    inp = _f32(inp)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    lib().oracle_fps(_int(b), _int(n), _int(npoint), _pf(inp), _pi(out), _int(threads))
    return out


def gather(inp, idx) -> np.ndarray:
    inp, idx = _f32(inp), _i32(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.zeros((b, m, 3), np.float32)
    lib().oracle_gather(_int(b), _int(n), _int(m), _pf(inp), _pi(idx), _pf(out))
    return out


def gather_grad(inp_shape, idx, out_g) -> np.ndarray:
    idx, out_g = _i32(idx), _f32(out_g)
    b, n, _ = inp_shape
    m = idx.shape[1]
    inp_g = np.zeros((b, n, 3), np.float32)
    lib().oracle_gather_grad(_int(b), _int(n), _int(m), _pf(out_g), _pi(idx), _pf(inp_g))
    return inp_g


def nn_distance(xyz1, xyz2, mode: str = "gpu", threads: int = 1):
    """nn_distance(xyz1, xyz2) -> (dist1, idx1, dist2, idx2). mode 'gpu' = FMA order of
    tf_nndistance_g.cu, 'cpu' = nnsearch order of tf_nndistance.cpp."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((b, n), np.float32); i1 = np.zeros((b, n), np.int32)
    d2 = np.zeros((b, m), np.float32); i2 = np.zeros((b, m), np.int32)
    lib().oracle_nn_distance(_int(b), _int(n), _pf(xyz1), _int(m), _pf(xyz2), _pf(d1), _pi(i1), _pf(d2), _pi(i2),
                             _int({"gpu": 0, "cpu": 1}[mode]), _int(threads))
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    gd1, gd2, idx1, idx2 = _f32(grad_dist1), _f32(grad_dist2), _i32(idx1), _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros((b, n, 3), np.float32); g2 = np.zeros((b, m, 3), np.float32)
    lib().oracle_nn_distance_grad(_int(b), _int(n), _pf(xyz1), _int(m), _pf(xyz2), _pf(gd1), _pi(idx1), _pf(gd2),
                                  _pi(idx2), _pf(g1), _pf(g2))
    return g1, g2


def cumsum(inp) -> np.ndarray:
    inp = _f32(inp)
    b, n = inp.shape
    out = np.zeros((b, n), np.float32)
    lib().oracle_cumsum(_int(b), _int(n), _pf(inp), _pf(out))
    return out


def prob_sample(inp_p, inp_r) -> np.ndarray:
    inp_p, inp_r = _f32(inp_p), _f32(inp_r)
    b, n = inp_p.shape
    m = inp_r.shape[1]
    temp = np.zeros((b, n), np.float32)
    out = np.zeros((b, m), np.int32)
    lib().oracle_prob_sample(_int(b), _int(n), _int(m), _pf(inp_p), _pf(inp_r), _pf(temp), _pi(out))
    return out


# ------------------------------------------------------------------ the reference's own code
class RefError(RuntimeError):
    """The reference op failed an OP_REQUIRES check (message = the reference's own)."""


def ref_cpu_nn_distance(xyz1, xyz2):
    """Runs NnDistanceOp::Compute (tf_nndistance.cpp:45-84) — the reference's CPU kernel, unmodified."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    s1 = (ctypes.c_longlong * max(xyz1.ndim, 1))(*xyz1.shape)
    s2 = (ctypes.c_longlong * max(xyz2.ndim, 1))(*xyz2.shape)
    b = xyz1.shape[0]
    n = xyz1.shape[1] if xyz1.ndim > 1 else 1
    m = xyz2.shape[1] if xyz2.ndim > 1 else 1
    d1 = np.zeros((b, n), np.float32); i1 = np.zeros((b, n), np.int32)
    d2 = np.zeros((b, m), np.float32); i2 = np.zeros((b, m), np.int32)
    rc = ref().ref_cpu_nn_distance(_pf(xyz1), _int(xyz1.ndim), s1, _pf(xyz2), _int(xyz2.ndim), s2, _pf(d1), _pi(i1),
                                   _pf(d2), _pi(i2))
    if rc != 0:
        raise RefError(ref().ref_last_error().decode())
    return d1, i1, d2, i2


def ref_cpu_nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    """Runs NnDistanceGradOp::Compute (tf_nndistance.cpp:86-165)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    gd1, gd2, idx1, idx2 = _f32(grad_dist1), _f32(grad_dist2), _i32(idx1), _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros((b, n, 3), np.float32); g2 = np.zeros((b, m, 3), np.float32)
    rc = ref().ref_cpu_nn_distance_grad(_int(b), _int(n), _pf(xyz1), _int(m), _pf(xyz2), _pf(gd1), _pi(idx1),
                                        _pf(gd2), _pi(idx2), _pf(g1), _pf(g2))
    if rc != 0:
        raise RefError(ref().ref_last_error().decode())
    return g1, g2


def _dp(t):
    return ctypes.c_void_p(t.data_ptr())


def ref_gpu_nn_distance(xyz1, xyz2):
    """Reference NmDistanceKernelLauncher rebuilt for sm_100a, on torch CUDA tensors."""
    import torch
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = torch.empty(b, n, device=xyz1.device); i1 = torch.empty(b, n, device=xyz1.device, dtype=torch.int32)
    d2 = torch.empty(b, m, device=xyz1.device); i2 = torch.empty(b, m, device=xyz1.device, dtype=torch.int32)
    torch.cuda.synchronize()
    ref().ref_gpu_nn_distance(_int(b), _int(n), _dp(xyz1), _int(m), _dp(xyz2), _dp(d1), _dp(i1), _dp(d2), _dp(i2))
    torch.cuda.synchronize()
    return d1, i1, d2, i2


def ref_gpu_nn_distance_grad(xyz1, xyz2, gd1, idx1, gd2, idx2):
    import torch
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = torch.empty(b, n, 3, device=xyz1.device); g2 = torch.empty(b, m, 3, device=xyz1.device)
    torch.cuda.synchronize()
    ref().ref_gpu_nn_distance_grad(_int(b), _int(n), _dp(xyz1), _int(m), _dp(xyz2), _dp(gd1), _dp(idx1), _dp(gd2),
                                   _dp(idx2), _dp(g1), _dp(g2))
    torch.cuda.synchronize()
    return g1, g2


def ref_gpu_fps(inp, npoint: int):
    import torch
    b, n, _ = inp.shape
    temp = torch.empty(32, n, device=inp.device)
    out = torch.empty(b, npoint, device=inp.device, dtype=torch.int32)
    torch.cuda.synchronize()
    ref().ref_gpu_fps(_int(b), _int(n), _int(npoint), _dp(inp), _dp(temp), _dp(out))
    torch.cuda.synchronize()
    return out


def ref_gpu_gather(inp, idx):
    import torch
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = torch.empty(b, m, 3, device=inp.device)
    torch.cuda.synchronize()
    ref().ref_gpu_gather(_int(b), _int(n), _int(m), _dp(inp), _dp(idx), _dp(out))
    torch.cuda.synchronize()
    return out


def ref_gpu_gather_grad(inp, idx, out_g):
    import torch
    b, n, _ = inp.shape
    m = idx.shape[1]
    inp_g = torch.zeros(b, n, 3, device=inp.device)
    torch.cuda.synchronize()
    ref().ref_gpu_gather_grad(_int(b), _int(n), _int(m), _dp(out_g), _dp(idx), _dp(inp_g))
    torch.cuda.synchronize()
    return inp_g


def ref_gpu_prob_sample(inp_p, inp_r):
    import torch
    b, n = inp_p.shape
    m = inp_r.shape[1]
    temp = torch.empty(b, n, device=inp_p.device)
    out = torch.empty(b, m, device=inp_p.device, dtype=torch.int32)
    torch.cuda.synchronize()
    ref().ref_gpu_prob_sample(_int(b), _int(n), _int(m), _dp(inp_p), _dp(inp_r), _dp(temp), _dp(out))
    torch.cuda.synchronize()
    return out
