"""Resnet ("standard" backbone) -- mirror of the reference's ``models/Resnet.py:43-99``:
3x3 stride-2 stem, residual blocks that pool while H > S, 3x3 pad-1 head."""
from __future__ import annotations

from .PoolResnet import GridBackbone


class Resnet(GridBackbone):
    def __init__(self, filters, input_shape, num_of_patches, num_of_residual_blocks=10, probability_threshold=0.5,
                 iou_threshold=0.5, pretrained=False, output_kernel_size=3):
        super().__init__(filters, input_shape, num_of_patches=num_of_patches,
                         probability_threshold=probability_threshold, iou_threshold=iou_threshold)
        self.pretrained = pretrained
        S = self.num_of_patches
        self._build(filters, input_shape, num_of_residual_blocks, 3, 2, 1,          # Resnet.py:64-70
                    output_kernel_size, 1,                                           # Resnet.py:77-83
                    pool_rule=lambda h: h > S,                                       # Resnet.py:38
                    block_patches=S)
