"""Multi-scale grid-cell assignment (boxes -> [4774,5] SSD target) on the GPU.

Mirror of ``WIDERFaceDatasetSSD.convert_bbx_to_feature_map`` + the per-scale concatenation of ``__getitem__``
(reference datasets/WIDERFace/dataset_ssd.py:36-76,134-139).  File I/O and augmentation are out of scope.
"""
from __future__ import annotations

from typing import Sequence

import torch

from ... import ops

PATCH_SIZES = (60, 30, 15, 7)


def convert_bbx_to_feature_maps_batch(boxes: Sequence[torch.Tensor], img_size, patch_sizes=PATCH_SIZES, device=None):
    """Ragged list of ``[K_i,5]`` (1,x,y,w,h) pixel boxes -> ``[B,P,5]`` f32 on the GPU, one launch."""
    device = device or (boxes[0].device if boxes[0].is_cuda else torch.device("cuda"))
    width, height = img_size
    offs = [0]
    for b in boxes:
        offs.append(offs[-1] + (int(b.shape[0]) if b.dim() == 2 else 0))
    flat = (torch.cat([b.reshape(-1, 5).float() for b in boxes if b.dim() == 2]) if offs[-1]
            else torch.zeros((1, 5)))
    flat = flat.to(device).contiguous()
    offsets = torch.tensor(offs, dtype=torch.int32).to(device)
    P = sum(ps * ps for ps in patch_sizes)
    out = torch.empty((len(boxes), P, 5), dtype=torch.float32, device=device)
    ops.ssd_grid_encode(flat, offsets, patch_sizes, width, height, out)
    return out


class WIDERFaceDatasetSSD:
    """Only the encoder of the reference class (dataset_ssd.py:14-76); same constructor arguments."""

    def __init__(self, data_dir, num_of_patches, input_shape, targets=None, split: str = "train", transform=None):
        self.data_dir = data_dir
        self.transform = transform
        self.targets = targets
        self.num_of_patches = num_of_patches
        self.input_shape = input_shape
        self.patch_sizes = PATCH_SIZES

    def convert_bbx_to_feature_map(self, bbx, img_size, patch_size):
        """One scale, reference layout ``[5,ps,ps]`` (dataset_ssd.py:36-76)."""
        fm = convert_bbx_to_feature_maps_batch([bbx], img_size, (patch_size,))[0]
        return fm.reshape(patch_size, patch_size, 5).permute(2, 0, 1).contiguous()

    def encode(self, bbx, img_size=None):
        """All scales concatenated, ``[P,5]`` (dataset_ssd.py:134-139)."""
        return convert_bbx_to_feature_maps_batch([bbx], img_size or self.input_shape, self.patch_sizes)[0]
