"""SSD loss (hard-negative mining + clamped BCE + smooth L1), forward + backward in one kernel.

Mirror of the reference's ``losses/SSDLoss.py:27-86``: ``ssd_loss(confidence[B,P], predicted_locations[B,P,4],
labels[B,P], gt_locations[B,P,4], neg_pos_ratio) -> 0-d tensor`` differentiable w.r.t. ``confidence`` and
``predicted_locations`` (the caller is models/ModelMetaSSD.py:175).  The two full sorts of the reference's mining
step are replaced by a per-row radix select in shared memory (``fd_ssd_loss``).
"""
from __future__ import annotations

import math

import torch

from .. import ops


class _SsdLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, confidence, predicted_locations, labels, gt_locations, neg_pos_ratio, num_pos_reduce):
        c = confidence.detach().float().contiguous()
        l = predicted_locations.detach().float().contiguous()
        lb = labels.detach().float().contiguous()
        g = gt_locations.detach().float().contiguous()
        B, P = c.shape
        sums = torch.empty((B, 2), dtype=torch.float32, device=c.device)
        npos = torch.empty((B,), dtype=torch.int32, device=c.device)
        dconf, dloc = torch.empty_like(c), torch.empty_like(l)
        ops.ssd_loss(c, l, lb, g, int(neg_pos_ratio), sums, npos, None, dconf, dloc)
        n = npos.sum().float()                       # SSDLoss.py:85: positives of the WHOLE batch
        if num_pos_reduce is not None:               # data parallel: one integer all-reduce
            n = num_pos_reduce(n)
        ctx.save_for_backward(dconf, dloc, n)
        ctx.dtypes = (confidence.dtype, predicted_locations.dtype)
        return (sums[:, 1].sum() + sums[:, 0].sum()) / n

    @staticmethod
    def backward(ctx, dloss):
        dconf, dloc, n = ctx.saved_tensors
        s = dloss / n
        return (dconf * s).to(ctx.dtypes[0]), (dloc * s).to(ctx.dtypes[1]), None, None, None, None


def ssd_loss(confidence, predicted_locations, labels, gt_locations, neg_pos_ratio, num_pos_reduce=None):
    """Reference signature (losses/SSDLoss.py:57).  ``num_pos_reduce`` (optional callable) sums the positive
    count across data-parallel ranks."""
    return _SsdLossFn.apply(confidence, predicted_locations, labels, gt_locations, neg_pos_ratio, num_pos_reduce)


@torch.no_grad()
def hard_negative_mining(loss, labels, neg_pos_ratio):
    """Reference signature (losses/SSDLoss.py:27-53): ranks the negatives by the SUPPLIED ``loss`` (descending) and
    returns the bool mask; like the reference it also overwrites ``loss[labels > 0]`` with ``-inf`` in place (:45).
    The kernel selects on the raw key bits in ascending order, so ``loss`` is mapped to a key that is ascending in
    descending loss with an exact, order-preserving bit transform (no ``exp``/``log`` round trip that could merge or
    reorder near-tied losses): flip all bits of a negative float, the sign bit of a non-negative one, then invert."""
    lf = loss.detach().float().contiguous()
    B, P = lf.shape
    lb = labels.detach().float().contiguous()
    loss[labels > 0] = -math.inf                                  # SSDLoss.py:45 (in-place side effect)
    bits = lf.view(torch.int32)
    asc = torch.where(bits < 0, ~bits, bits ^ (-0x80000000))      # unsigned-ascending in ascending loss
    key = (~asc).view(torch.float32)                              # unsigned-ascending in DESCENDING loss
    z4 = torch.zeros((B, P, 4), dtype=torch.float32, device=lf.device)
    sums = torch.empty((B, 2), dtype=torch.float32, device=lf.device)
    npos = torch.empty((B,), dtype=torch.int32, device=lf.device)
    mask = torch.empty((B, P), dtype=torch.uint8, device=lf.device)
    ops.ssd_loss(key, z4, lb, z4, int(neg_pos_ratio), sums, npos, mask, None, None)
    return mask.bool()
