"""Sampling ops — drop-in for the reference's ``tf_ops/sampling/tf_sampling.py``:
``prob_sample`` (:13-22), ``gather_point`` (:29-37, gradient :43-47) and
``farthest_point_sample`` (:48-57, no gradient).

Same names, argument order, dtypes and return shapes; tensors are torch CUDA tensors, kernels are
the sm_100a ones behind the ``caae_*`` C ABI.  Like the reference (``REGISTER_KERNEL_BUILDER ...
DEVICE_GPU`` only, tf_sampling.cpp:92,123,148,178) there is no CPU kernel.
"""
from __future__ import annotations

import torch

from ... import _capi
from ..._capi import InvalidArgumentError

__all__ = ["prob_sample", "gather_point", "gather_point_grad", "farthest_point_sample",
           "farthest_point_sample_gather"]


# ------------------------------------------------------------------ farthest_point_sample
def _fps_impl(inp: torch.Tensor, npoint: int, want_xyz: bool):
    b, n, _ = inp.shape
    inp = inp.contiguous()
    out = torch.empty((b, npoint), dtype=torch.int32, device=inp.device)
    out_xyz = torch.empty((b, npoint, 3), dtype=torch.float32, device=inp.device) if want_xyz else None
    nbytes = _capi.lib().caae_fps_scratch_bytes(b, n)
    temp = torch.empty(nbytes // 4, dtype=torch.float32, device=inp.device) if nbytes else None
    with torch.cuda.device(inp.device):
        _capi.check(_capi.lib().caae_fps_gather(b, n, npoint, _capi.ptr(inp), _capi.ptr(temp), _capi.ptr(out),
                                                _capi.ptr(out_xyz), _capi.stream_of(inp)), "caae_fps")
    return out, out_xyz


@torch.library.custom_op("cloudaae::farthest_point_sample", mutates_args=(), device_types="cuda")
def _fps_op(inp: torch.Tensor, npoint: int) -> torch.Tensor:
    return _fps_impl(inp, npoint, False)[0]


@_fps_op.register_fake
def _(inp, npoint):
    return inp.new_empty((inp.shape[0], npoint), dtype=torch.int32)


@torch.library.custom_op("cloudaae::farthest_point_sample_gather", mutates_args=(), device_types="cuda")
def _fps_gather_op(inp: torch.Tensor, npoint: int) -> tuple[torch.Tensor, torch.Tensor]:
    out, out_xyz = _fps_impl(inp, npoint, True)
    return out, out_xyz


@_fps_gather_op.register_fake
def _(inp, npoint):
    return (inp.new_empty((inp.shape[0], npoint), dtype=torch.int32), inp.new_empty((inp.shape[0], npoint, 3)))


def _check_fps(npoint, inp):
    if not isinstance(npoint, int) or npoint <= 0:
        raise InvalidArgumentError("FarthestPointSample expects positive npoint")
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise InvalidArgumentError("FarthestPointSample expects (batch_size,num_points,3) inp shape")
    if inp.dtype != torch.float32:
        raise InvalidArgumentError("FarthestPointSample expects float32 inp")
    if inp.shape[0] > 0 and inp.shape[1] == 0:
        raise InvalidArgumentError("FarthestPointSample cannot sample from an empty cloud")
    _capi.require_cuda(inp, "farthest_point_sample")


def farthest_point_sample(npoint: int, inp: torch.Tensor) -> torch.Tensor:
    """input: int32 npoint; inp (batch_size, ndataset, 3) float32.  returns (batch_size, npoint) int32.

    Seed index 0; ties broken exactly as the reference kernel does.  Not differentiable.
    """
    _check_fps(npoint, inp)
    return _fps_op(inp.detach(), npoint)


def farthest_point_sample_gather(npoint: int, inp: torch.Tensor):
    """Fused ``gather_point(inp, farthest_point_sample(npoint, inp))`` — the composition the
    evaluation graph runs (evaluate_cloudAAE_ycbv.py:450).  Returns (idx, xyz); xyz carries the
    gather gradient."""
    _check_fps(npoint, inp)
    idx, xyz = _fps_gather_op(inp.detach(), npoint)
    if inp.requires_grad:
        xyz = gather_point(inp, idx)
    return idx, xyz


# ------------------------------------------------------------------ gather_point (+grad)
@torch.library.custom_op("cloudaae::gather_point", mutates_args=(), device_types="cuda")
def _gather_op(inp: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    b, n, _ = inp.shape
    m = idx.shape[1]
    inp, idx = inp.contiguous(), idx.contiguous()
    out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
    with torch.cuda.device(inp.device):
        _capi.check(_capi.lib().caae_gather(b, n, m, _capi.ptr(inp), _capi.ptr(idx), _capi.ptr(out),
                                            _capi.stream_of(inp)), "caae_gather")
    return out


@_gather_op.register_fake
def _(inp, idx):
    return inp.new_empty((inp.shape[0], idx.shape[1], 3))


@torch.library.custom_op("cloudaae::gather_point_grad", mutates_args=(), device_types="cuda")
def _gather_grad_op(inp: torch.Tensor, idx: torch.Tensor, out_g: torch.Tensor) -> torch.Tensor:
    b, n, _ = inp.shape
    m = idx.shape[1]
    idx, out_g = idx.contiguous(), out_g.contiguous()
    inp_g = torch.empty((b, n, 3), dtype=torch.float32, device=inp.device)
    with torch.cuda.device(inp.device):
        _capi.check(_capi.lib().caae_gather_grad(b, n, m, _capi.ptr(out_g), _capi.ptr(idx), _capi.ptr(inp_g),
                                                 _capi.stream_of(inp)), "caae_gather_grad")
    return inp_g


@_gather_grad_op.register_fake
def _(inp, idx, out_g):
    return torch.empty_like(inp)


def _gather_setup(ctx, inputs, output):
    inp, idx = inputs
    ctx.save_for_backward(inp, idx)


def _gather_backward(ctx, out_g):
    inp, idx = ctx.saved_tensors
    return _gather_grad_op(inp, idx, out_g), None  # tf_sampling.py:43-47


_gather_op.register_autograd(_gather_backward, setup_context=_gather_setup)


def _check_gather(op, inp, idx):
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise InvalidArgumentError(f"{op} expects (batch_size,num_points,3) inp shape")
    if idx.dim() != 2 or idx.shape[0] != inp.shape[0]:
        raise InvalidArgumentError(f"{op} expects (batch_size,num_result) idx shape")
    if inp.dtype != torch.float32 or idx.dtype != torch.int32:
        raise InvalidArgumentError(f"{op} expects float32 inp and int32 idx")
    if idx.shape[1] > 0 and inp.shape[1] == 0 and inp.shape[0] > 0:
        raise InvalidArgumentError(f"{op} cannot gather from an empty cloud")


def gather_point(inp: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """input: inp (batch_size, ndataset, 3) float32; idx (batch_size, npoints) int32.
    returns (batch_size, npoints, 3) float32."""
    _check_gather("GatherPoint", inp, idx)
    _capi.require_cuda(inp, "gather_point")
    return _gather_op(inp, idx)


def gather_point_grad(inp: torch.Tensor, idx: torch.Tensor, out_g: torch.Tensor) -> torch.Tensor:
    """``sampling_module.gather_point_grad`` (GatherPointGrad op, tf_sampling.cpp:151-178)."""
    _check_gather("GatherPointGradGpuOp", inp, idx)
    if out_g.dim() != 3 or tuple(out_g.shape) != (inp.shape[0], idx.shape[1], 3):
        raise InvalidArgumentError("GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape")
    _capi.require_cuda(inp, "gather_point_grad")
    return _gather_grad_op(inp, idx, out_g.float())


# ------------------------------------------------------------------ prob_sample
@torch.library.custom_op("cloudaae::prob_sample", mutates_args=(), device_types="cuda")
def _prob_sample_op(inp: torch.Tensor, inpr: torch.Tensor) -> torch.Tensor:
    b, n = inp.shape
    m = inpr.shape[1]
    inp, inpr = inp.contiguous(), inpr.contiguous()
    temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
    out = torch.empty((b, m), dtype=torch.int32, device=inp.device)
    with torch.cuda.device(inp.device):
        _capi.check(_capi.lib().caae_prob_sample(b, n, m, _capi.ptr(inp), _capi.ptr(inpr), _capi.ptr(temp),
                                                 _capi.ptr(out), _capi.stream_of(inp)), "caae_prob_sample")
    return out


@_prob_sample_op.register_fake
def _(inp, inpr):
    return inp.new_empty((inp.shape[0], inpr.shape[1]), dtype=torch.int32)


def prob_sample(inp: torch.Tensor, inpr: torch.Tensor) -> torch.Tensor:
    """input: inp (batch_size, ncategory) float32 weights; inpr (batch_size, npoints) float32 in [0,1).
    returns (batch_size, npoints) int32 — category drawn with probability proportional to inp."""
    if inp.dim() != 2:
        raise InvalidArgumentError("ProbSample expects (batch_size,num_choices) inp shape")
    if inpr.dim() != 2 or inpr.shape[0] != inp.shape[0]:
        raise InvalidArgumentError("ProbSample expects (batch_size,num_points) inpr shape")
    if inp.dtype != torch.float32 or inpr.dtype != torch.float32:
        raise InvalidArgumentError("ProbSample expects float32 inputs")
    if inp.shape[0] > 0 and inpr.shape[1] > 0 and inp.shape[1] == 0:
        raise InvalidArgumentError("ProbSample expects at least one category")
    _capi.require_cuda(inp, "prob_sample")
    return _prob_sample_op(inp.detach(), inpr.detach())
