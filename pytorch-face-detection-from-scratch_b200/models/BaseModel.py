"""nn.Module base of the detectors -- mirror of the reference's ``models/BaseModel.py:11-71``
(same constructor, attributes and methods) with decode + NMS running as one batched CUDA kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..datasets.utils import ReduceBoundingBoxes


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
        size = tuple(self.input_shape[1:])
        if tuple(x.shape[-2:]) == size:
            return x
        from torchvision.transforms import transforms
        return transforms.Resize(size=size)(x)

    @torch.no_grad()
    def predict(self, x, probability_threshold=0.5, iou_threshold=0.5):
        """BaseModel.py:56-71."""
        self.reduce_bounding_boxes = ReduceBoundingBoxes(
            probability_threshold=probability_threshold,
            iou_threshold=iou_threshold,
            input_shape=self.input_shape,
            num_of_patches=self.num_of_patches,
        )
        x = self._resize(x)
        x = x / 255.0
        image = x
        if len(x.shape) == 3:
            x = x.reshape(1, *x.shape)
        x = self(x)
        bbxs = self.non_max_suppression(x)
        return image, bbxs[0]
