"""nn.Module base of the detectors -- mirror of the reference's ``models/BaseModel.py:11-71``
(same constructor, attributes and methods) with decode + NMS running as one batched CUDA kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..datasets.utils import ReduceBoundingBoxes


def resize_to(x: torch.Tensor, size) -> torch.Tensor:
    """``transforms.Resize(size)(x)`` of models/BaseModel.py:64 / PoolResnet.py:94 as one CUDA kernel (``fd_resize_bilinear``):
    bilinear, align_corners=False, no antialias (the pinned torchvision 0.11.2 semantics, which are also what the
    official TorchScript archives execute); uint8 in -> uint8 out (round half to even), float in -> float32 out.
    ``x``: ``[3,h,w]`` or ``[B,3,h,w]``.  Same size -> returned unchanged, like the reference's no-op resize."""
    size = tuple(int(s) for s in size)
    if tuple(x.shape[-2:]) == size:
        return x
    if not x.is_cuda:
        raise RuntimeError("fd_b200 resize runs on CUDA tensors only (no CPU fallback)")
    squeeze = x.dim() == 3
    xb = x.unsqueeze(0) if squeeze else x
    if xb.dtype != torch.uint8:
        xb = xb.float()
    xb = xb.contiguous()
    out = torch.empty((xb.shape[0], xb.shape[1], size[0], size[1]), dtype=xb.dtype, device=xb.device)
    ops.resize_bilinear(xb, out)
    return out[0] if squeeze else out


class BaseModel(nn.Module):
    def __init__(self, filters, input_shape, num_of_patches, probability_threshold=0.5, iou_threshold=0.5):
        super().__init__()
        self.input_shape = input_shape
        self.num_of_patches = num_of_patches
        assert (
            input_shape[1] % num_of_patches == 0 and input_shape[2] % num_of_patches == 0
        ), f"Input shape {input_shape} cannot be divided into {num_of_patches} patches"   # BaseModel.py:23-26
        self.probability_threshold = probability_threshold
        self.iou_threshold = iou_threshold
        self.reduce_bounding_boxes = ReduceBoundingBoxes(
            probability_threshold=probability_threshold,
            iou_threshold=iou_threshold,
            input_shape=self.input_shape,
            num_of_patches=self.num_of_patches,
        )

    def summary(self):
        """BaseModel.py:36-45 (torchinfo is not a dependency here: prints the parameter count)."""
        if self.input_shape is None:
            raise Exception("Please set 'input_shape'")
        n = sum(p.numel() for p in self.parameters())
        print(f"{type(self).__name__}: {n} parameters, input {tuple(self.input_shape)}")

    def non_max_suppression(self, x):
        """BaseModel.py:47-51: 4-D input -> ragged tuple of [K_i,5]; 3-D -> one [K,5] tensor.
        The batch is decoded by ONE kernel launch when the reducer is ours; a user-replaced
        ``reduce_bounding_boxes`` callable (pruner.py:33-38 does that) is honoured per image."""
        if len(x.shape) == 4:
            rb = self.reduce_bounding_boxes
            if isinstance(rb, ReduceBoundingBoxes) and x.is_cuda:
                boxes, counts = rb.batch_forward(x)
                return rb.batch_to_tuple(boxes, counts)
            return tuple([rb(x[i]) for i in range(x.shape[0])])
        return self.reduce_bounding_boxes(x)

    def single_non_max_suppression(self, x):
        return self.reduce_bounding_boxes(x)

    def _resize(self, x):
        return resize_to(x, tuple(self.input_shape[1:]))

    def _to_model_device(self, x):
        """Host images are copied to the parameters' device (plumbing; the arithmetic has no CPU path)."""
        p = next(self.parameters(), None)
        if p is not None and p.is_cuda and not x.is_cuda:
            return x.to(p.device, non_blocking=True)
        return x

    def to_torchscript(self, file_path=None, method="script", example_inputs=None, **kwargs):
        """``torch.jit.script(model)`` of demo_scripts/convert_checkpoint_to_scripted_model.py:51 / LightningModule's
        ``to_torchscript`` (train_model.py:61): a ScriptModule with the same ``forward(x, predict)`` that calls the
        registered ``fd_b200`` operators (export.py); saved to ``file_path`` when given."""
        from ..export import to_torchscript
        return to_torchscript(self, file_path)

    @torch.no_grad()
    def predict(self, x, probability_threshold=0.5, iou_threshold=0.5):
        """BaseModel.py:56-71."""
        self.reduce_bounding_boxes = ReduceBoundingBoxes(
            probability_threshold=probability_threshold,
            iou_threshold=iou_threshold,
            input_shape=self.input_shape,
            num_of_patches=self.num_of_patches,
        )
        x = self._resize(self._to_model_device(x))
        x = x / 255.0
        image = x
        if len(x.shape) == 3:
            x = x.reshape(1, *x.shape)
        x = self(x)
        bbxs = self.non_max_suppression(x)
        return image, bbxs[0]
