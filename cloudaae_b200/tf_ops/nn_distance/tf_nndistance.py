"""nn_distance — chamfer nearest-neighbour distance, drop-in for the reference's
``tf_ops/nn_distance/tf_nndistance.py`` (nn_distance :14-24, registered gradient :31-37).

Same names, argument order, dtypes and return tuple; tensors are torch CUDA tensors and the
kernels are the hand-written sm_100a ones behind ``caae_nn_distance`` / ``caae_nn_distance_grad``.
Shape violations raise :class:`InvalidArgumentError` with the reference op's own message
(tf_nndistance.cpp:175-182, 222-233).
"""
from __future__ import annotations

import torch

from ... import _capi
from ..._capi import InvalidArgumentError

__all__ = ["nn_distance", "nn_distance_grad"]


def _check_pair(op: str, xyz1: torch.Tensor, xyz2: torch.Tensor) -> None:
    if xyz1.dim() != 3:
        raise InvalidArgumentError(f"{op} requires xyz1 be of shape (batch,#points,3)")
    if xyz1.shape[2] != 3:
        raise InvalidArgumentError(f"{op} only accepts 3d point set xyz1")
    if xyz2.dim() != 3:
        raise InvalidArgumentError(f"{op} requires xyz2 be of shape (batch,#points,3)")
    if xyz2.shape[2] != 3:
        raise InvalidArgumentError(f"{op} only accepts 3d point set xyz2")
    if xyz2.shape[0] != xyz1.shape[0]:
        raise InvalidArgumentError(f"{op} expects xyz1 and xyz2 have same batch size")
    if xyz1.dtype != torch.float32 or xyz2.dtype != torch.float32:
        raise InvalidArgumentError(f"{op} expects float32 inputs")


@torch.library.custom_op("cloudaae::nn_distance", mutates_args=(), device_types="cuda")
def _nn_distance_op(xyz1: torch.Tensor, xyz2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    xyz1 = xyz1.contiguous()
    xyz2 = xyz2.contiguous()
    dist1 = torch.empty((b, n), dtype=torch.float32, device=xyz1.device)
    idx1 = torch.empty((b, n), dtype=torch.int32, device=xyz1.device)
    dist2 = torch.empty((b, m), dtype=torch.float32, device=xyz1.device)
    idx2 = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        _capi.check(_capi.lib().caae_nn_distance(b, n, _capi.ptr(xyz1), m, _capi.ptr(xyz2), _capi.ptr(dist1),
                                                 _capi.ptr(idx1), _capi.ptr(dist2), _capi.ptr(idx2),
                                                 _capi.stream_of(xyz1)), "caae_nn_distance")
    return dist1, idx1, dist2, idx2


@_nn_distance_op.register_fake
def _(xyz1, xyz2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    return (xyz1.new_empty((b, n)), xyz1.new_empty((b, n), dtype=torch.int32),
            xyz1.new_empty((b, m)), xyz1.new_empty((b, m), dtype=torch.int32))


@torch.library.custom_op("cloudaae::nn_distance_grad", mutates_args=(), device_types="cuda")
def _nn_distance_grad_op(xyz1: torch.Tensor, xyz2: torch.Tensor, grad_dist1: torch.Tensor, idx1: torch.Tensor,
                         grad_dist2: torch.Tensor, idx2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
    grad_dist1, grad_dist2 = grad_dist1.contiguous(), grad_dist2.contiguous()
    idx1, idx2 = idx1.contiguous(), idx2.contiguous()
    grad_xyz1 = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    grad_xyz2 = torch.empty((b, m, 3), dtype=torch.float32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        _capi.check(_capi.lib().caae_nn_distance_grad(
            b, n, _capi.ptr(xyz1), m, _capi.ptr(xyz2), _capi.ptr(grad_dist1), _capi.ptr(idx1), _capi.ptr(grad_dist2),
            _capi.ptr(idx2), _capi.ptr(grad_xyz1), _capi.ptr(grad_xyz2), _capi.stream_of(xyz1)),
            "caae_nn_distance_grad")
    return grad_xyz1, grad_xyz2


@_nn_distance_grad_op.register_fake
def _(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    return torch.empty_like(xyz1), torch.empty_like(xyz2)


def _setup_context(ctx, inputs, output):
    xyz1, xyz2 = inputs
    _, idx1, _, idx2 = output
    ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
    ctx.set_materialize_grads(False)


def _backward(ctx, grad_dist1, grad_idx1, grad_dist2, grad_idx2):
    # tf_nndistance.py:31-37 — the index outputs carry no gradient.
    xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
    if grad_dist1 is None:
        grad_dist1 = torch.zeros(idx1.shape, dtype=torch.float32, device=xyz1.device)
    if grad_dist2 is None:
        grad_dist2 = torch.zeros(idx2.shape, dtype=torch.float32, device=xyz1.device)
    g1, g2 = _nn_distance_grad_op(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2)
    return g1, g2


_nn_distance_op.register_autograd(_backward, setup_context=_setup_context)


def nn_distance(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """Computes the distance of nearest neighbors for a pair of point clouds.

    input:  xyz1 (batch_size,#points_1,3), xyz2 (batch_size,#points_2,3), float32
    output: dist1 (batch_size,#points_1) squared distance from first to second,
            idx1  (batch_size,#points_1) int32 nearest neighbour in the second cloud,
            dist2 (batch_size,#points_2), idx2 (batch_size,#points_2) — the other direction.
    """
    _check_pair("NnDistance", xyz1, xyz2)
    _capi.require_cuda(xyz1, "nn_distance")
    _capi.require_cuda(xyz2, "nn_distance")
    return _nn_distance_op(xyz1, xyz2)


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    """``nn_distance_module.nn_distance_grad`` (NnDistanceGrad op, tf_nndistance.cpp:209-254)."""
    op = "NnDistanceGrad"
    _check_pair(op, xyz1, xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if tuple(grad_dist1.shape) != (b, n):
        raise InvalidArgumentError(f"{op} requires grad_dist1 be of shape(batch,#points)")
    if tuple(idx1.shape) != (b, n):
        raise InvalidArgumentError(f"{op} requires idx1 be of shape(batch,#points)")
    if tuple(grad_dist2.shape) != (b, m):
        raise InvalidArgumentError(f"{op} requires grad_dist2 be of shape(batch,#points)")
    if tuple(idx2.shape) != (b, m):
        raise InvalidArgumentError(f"{op} requires idx2 be of shape(batch,#points)")
    if idx1.dtype != torch.int32 or idx2.dtype != torch.int32:
        raise InvalidArgumentError(f"{op} expects int32 indices")
    _capi.require_cuda(xyz1, "nn_distance_grad")
    return _nn_distance_grad_op(xyz1, xyz2, grad_dist1.float(), idx1, grad_dist2.float(), idx2)
