"""Shared autograd node over ``caae_pose_losses`` (rotation in float64, translation in float32)."""
from __future__ import annotations

import torch

from .. import _capi


class _PoseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rot_pred, axag_label, trans_pred, trans_label):
        b = rot_pred.shape[0]
        dev = rot_pred.device
        f32 = dict(dtype=torch.float32, device=dev)
        per_rot = torch.empty(b, dtype=torch.float64, device=dev)
        per_trans = torch.empty(b, **f32)
        d_rot = torch.empty(b, 3, **f32)
        d_trans = torch.empty(b, 3, **f32)
        zero = torch.zeros(b, 3, **f32)
        rp, al = rot_pred.detach().float().contiguous(), axag_label.detach().float().contiguous()
        tp, tl = trans_pred.detach().float().contiguous(), trans_label.detach().float().contiguous()
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().caae_pose_losses(b, rp.data_ptr(), al.data_ptr(), tp.data_ptr(), zero.data_ptr(),
                                                     tl.data_ptr(), 1.0, 1.0, per_rot.data_ptr(), per_trans.data_ptr(),
                                                     d_rot.data_ptr(), d_trans.data_ptr(), None,
                                                     torch.cuda.current_stream(dev).cuda_stream), "caae_pose_losses")
        ctx.save_for_backward(d_rot, d_trans)
        ctx.rot_dtype, ctx.trans_dtype = rot_pred.dtype, trans_pred.dtype
        return per_rot, per_trans

    @staticmethod
    def backward(ctx, g_rot, g_trans):
        d_rot, d_trans = ctx.saved_tensors
        gr = None if g_rot is None else (d_rot.double() * g_rot[:, None].double()).to(ctx.rot_dtype)
        gt = None if g_trans is None else (d_trans * g_trans[:, None].float()).to(ctx.trans_dtype)
        return gr, None, gt, None


def pose_errors(rot_pred, axag_label, trans_pred, trans_label):
    for t, name in ((rot_pred, "pred"), (axag_label, "label"), (trans_pred, "pred"), (trans_label, "label")):
        if t.dim() != 2 or t.shape[1] != 3:
            raise _capi.InvalidArgumentError(f"{name} must be (batch_size, 3)")
    _capi.require_cuda(rot_pred, "pose losses")
    return _PoseLoss.apply(rot_pred, axag_label, trans_pred, trans_label)
