"""Decode + score threshold + NMS of the YOLO grid head on the GPU.

Mirror of the reference's ``datasets/utils.py`` (same class / function names, argument meaning and
return conventions) for the hot path: ``ReduceBoundingBoxes`` (reference datasets/utils.py:95-170).
One batched kernel (``fd_decode_nms``) replaces ~25 tiny ATen kernels, three ``torch.where`` host
syncs and ``torchvision.ops.nms`` per image.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class ReduceBoundingBoxes(nn.Module):
    """reference datasets/utils.py:95-170.

    ``__call__(x[5,S1,S2]) -> [K,5]`` rows (score, x, y, w, h), descending score, fresh writable
    tensor on x's device; no candidate -> shape (0,5) (utils.py:170).  ``input_shape`` is
    (C, width, height) exactly as the reference unpacks it (utils.py:107).
    """

    def __init__(self, probability_threshold: float = 0.9, iou_threshold: float = 0.5,
                 input_shape=(3, 320, 240), num_of_patches=40):
        super().__init__()
        self.probability_threshold = probability_threshold
        self.iou_threshold = iou_threshold
        self.input_shape = input_shape
        _, self.width, self.height = input_shape
        self.num_of_patches = num_of_patches
        self.x_patch_size = self.width / num_of_patches
        self.y_patch_size = self.height / num_of_patches

    @torch.no_grad()
    def batch_forward(self, x: torch.Tensor, return_cells: bool = False):
        """x ``[B,5,S1,S2]`` -> (boxes ``[B,S1*S2,5]``, counts ``[B]`` int32[, cells]) on the device,
        no host synchronisation."""
        x = x.detach().float().contiguous()
        B, _, S1, S2 = x.shape
        boxes = torch.empty((B, S1 * S2, 5), dtype=torch.float32, device=x.device)
        cells = torch.empty((B, S1 * S2), dtype=torch.int32, device=x.device) if return_cells else None
        counts = torch.empty((B,), dtype=torch.int32, device=x.device)
        ops.decode_nms(x, self.probability_threshold, self.iou_threshold, self.width, self.height,
                       self.num_of_patches, boxes, cells, counts)
        return (boxes, counts, cells) if return_cells else (boxes, counts)

    def batch_to_tuple(self, boxes, counts):
        """Ragged tuple of ``[K_i,5]`` tensors (one D2H copy of the counts)."""
        n = counts.tolist()
        return tuple(boxes[i, :n[i]].clone() for i in range(len(n)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Host tensors (the reference decodes CPU ground-truth maps in its dataset and in ``draw_bbx``) are copied to
        the current CUDA device, decoded by the same kernel, and the rows copied back -- there is no CPU arithmetic."""
        on_host = not x.is_cuda
        boxes, counts = self.batch_forward((x.cuda() if on_host else x).unsqueeze(0))
        k = int(counts.item())
        if k == 0:
            return torch.empty(0).reshape(0, 5)      # reference returns a CPU tensor here (utils.py:170)
        out = boxes[0, :k].clone()
        return out.cpu() if on_host else out


class ReduceSSDBoundingBoxes(nn.Module):
    """reference datasets/utils.py:8-92.  ``__call__(x[P,5]) -> [K,5]`` rows (score, x, y, w, h) in NMS keep
    order; ``P = sum(ps*ps)`` priors; ``with_priors`` applies the prior scaling of utils.py:59-64 (the reference
    uses it for encoded ground truth, dataset_ssd.py:142-145)."""

    def __init__(self, probability_threshold: float = 0.9, iou_threshold: float = 0.5, input_shape=(3, 320, 240),
                 patch_sizes=(60, 30, 15, 7), priors=None, with_priors=False):
        super().__init__()
        self.probability_threshold = probability_threshold
        self.iou_threshold = iou_threshold
        self.input_shape = input_shape
        _, self.width, self.height = input_shape
        self.patch_sizes = tuple(patch_sizes)
        self.with_priors = with_priors
        self.priors = priors            # kept for interface parity; the kernel derives the priors from patch_sizes

    @torch.no_grad()
    def batch_forward(self, x: torch.Tensor):
        """x ``[B,P,5]`` -> (boxes ``[B,P,5]``, counts ``[B]`` int32) on the device, no host synchronisation."""
        x = x.detach().float().contiguous()
        B, P, _ = x.shape
        assert P == sum(ps * ps for ps in self.patch_sizes), "prior count does not match patch_sizes"
        boxes = torch.empty((B, P, 5), dtype=torch.float32, device=x.device)
        counts = torch.empty((B,), dtype=torch.int32, device=x.device)
        ops.ssd_decode_nms(x, self.patch_sizes, self.probability_threshold, self.iou_threshold, self.width,
                           self.height, self.with_priors, boxes, counts)
        return boxes, counts

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        on_host = not x.is_cuda                      # see ReduceBoundingBoxes.forward
        boxes, counts = self.batch_forward((x.cuda() if on_host else x).unsqueeze(0))
        k = int(counts.item())
        if k == 0:
            return torch.empty(0).reshape(0, 5)      # utils.py:92
        out = boxes[0, :k].clone()
        return out.cpu() if on_host else out


def convert_bbx_to_xyxy(bbx):
    """reference datasets/utils.py:173-174."""
    return bbx[0], bbx[1], bbx[0] + bbx[2], bbx[1] + bbx[3]
