"""nn.Module base of the SSD detector -- mirror of the reference's ``models/BaseSSDModel.py:10-68`` (same constructor,
attributes and methods); decode + NMS of the ``[B,4774,5]`` head runs as one batched CUDA kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..datasets.utils import ReduceSSDBoundingBoxes


class BaseSSDModel(nn.Module):
    def __init__(self, filters, input_shape, probability_threshold=0.5, iou_threshold=0.5, priors=None):
        super().__init__()
        self.input_shape = input_shape
        self.probability_threshold = probability_threshold
        self.iou_threshold = iou_threshold

    def summary(self, *args, **kwargs):
        """BaseSSDModel.py:24-42 (torchinfo / ptflops are not dependencies here: prints the parameter count)."""
        if self.input_shape is None:
            raise Exception("Please set 'input_shape'")
        n = sum(p.numel() for p in self.parameters())
        print(f"{type(self).__name__}: {n} parameters, input {tuple(self.input_shape)}")

    def non_max_suppression(self, x):
        """BaseSSDModel.py:44-48: 3-D input ``[B,P,5]`` -> ragged tuple of ``[K_i,5]``; 2-D -> one ``[K,5]`` tensor."""
        if len(x.shape) == 3:
            rb = self.reduce_bounding_boxes
            if isinstance(rb, ReduceSSDBoundingBoxes) and x.is_cuda:
                boxes, counts = rb.batch_forward(x)
                n = counts.tolist()
                return tuple(boxes[i, :n[i]].clone() for i in range(len(n)))
            return tuple([rb(x[i]) for i in range(x.shape[0])])
        return self.reduce_bounding_boxes(x)

    def single_non_max_suppression(self, x):
        return self.reduce_bounding_boxes(x)

    def to_torchscript(self, file_path=None, method="script", example_inputs=None, **kwargs):
        from ..export import to_torchscript
        return to_torchscript(self, file_path)

    def _resize(self, x):
        from .BaseModel import resize_to
        return resize_to(x, tuple(self.input_shape[1:]))

    @torch.no_grad()
    def predict(self, x, probability_threshold=0.5, iou_threshold=0.5):
        """BaseSSDModel.py:53-68."""
        self.reduce_bounding_boxes = ReduceSSDBoundingBoxes(
            probability_threshold=probability_threshold,
            iou_threshold=iou_threshold,
            input_shape=self.input_shape,
            patch_sizes=self.patch_sizes,
        )
        x = self._resize(x)
        x = x / 255.0
        image = x
        if len(x.shape) == 3:
            x = x.reshape(1, *x.shape)
        x = self(x)
        bbxs = self.non_max_suppression(x)
        return image, bbxs[0]
