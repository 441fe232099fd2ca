"""YOLO sum-of-squares loss, forward + backward in one kernel.

Mirror of the reference's ``losses/YoloLoss.py:4-44``: ``yolo_loss(pred_fm[5,S,S], gt_fm[5,S,S])``
returns a 0-d tensor differentiable w.r.t. ``pred_fm``.  ``yolo_loss_batch`` is the same thing for a
whole batch in ONE launch (the reference's caller loops over the batch and sums,
models/ModelMeta.py:173-176).
"""
from __future__ import annotations

import torch

from .. import ops


class _YoloLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        p = pred.detach().float().contiguous()
        g = gt.detach().float().contiguous()
        loss = torch.empty((p.shape[0],), dtype=torch.float32, device=p.device)
        dpred = torch.empty_like(p)
        ops.yolo_loss(p, g, loss, None, dpred)      # gradient comes out of the same pass
        ctx.save_for_backward(dpred)
        ctx.in_dtype = pred.dtype
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dpred,) = ctx.saved_tensors
        return (dpred * dloss.view(-1, 1, 1, 1)).to(ctx.in_dtype), None


def yolo_loss_per_image(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """``[B,5,S1,S2]`` x2 -> ``[B]`` per-image losses."""
    return _YoloLossFn.apply(pred, gt)


def yolo_loss_batch(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Sum over the batch of the per-image losses (ModelMeta.py:173-176,215)."""
    return yolo_loss_per_image(pred, gt).sum()


def yolo_loss(pred_fm: torch.Tensor, gt_fm: torch.Tensor) -> torch.Tensor:
    """Reference signature (losses/YoloLoss.py:4): one image, 0-d result."""
    return yolo_loss_per_image(pred_fm.unsqueeze(0), gt_fm.unsqueeze(0))[0]
