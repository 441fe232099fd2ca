"""Grid-cell assignment (boxes -> [5,S,S] target map) on the GPU.

Mirror of ``WIDERFaceDataset.convert_bbx_to_feature_map`` (reference
datasets/WIDERFace/dataset.py:32-64).  File I/O and augmentation of the reference dataset class are
out of scope (synthetic batches); only the encoder is on the hot path.
"""
from __future__ import annotations

from typing import Sequence

import torch

from ... import ops


def convert_bbx_to_feature_map_batch(boxes: Sequence[torch.Tensor], num_of_patches: int, img_size, device=None):
    """Ragged list of ``[K_i,5]`` (1,x,y,w,h) boxes -> ``[B,5,S,S]`` f32 on the GPU, one launch."""
    device = device or (boxes[0].device if boxes[0].is_cuda else torch.device("cuda"))
    width, height = img_size
    offs = [0]
    for b in boxes:
        offs.append(offs[-1] + int(b.shape[0]))
    flat = torch.cat([b.reshape(-1, 5).float() for b in boxes]) if offs[-1] else torch.zeros((1, 5))
    flat = flat.to(device).contiguous()
    offsets = torch.tensor(offs, dtype=torch.int32).to(device)
    out = torch.empty((len(boxes), 5, num_of_patches, num_of_patches), dtype=torch.float32, device=device)
    ops.grid_encode(flat, offsets, num_of_patches, width, height, out)
    return out


class WIDERFaceDataset:
    """Only the encoder of the reference class (dataset.py:11-64); same constructor arguments."""

    def __init__(self, data_dir, num_of_patches, input_shape, targets=None, split: str = "train", transform=None):
        self.data_dir = data_dir
        self.transform = transform
        self.targets = targets
        self.num_of_patches = num_of_patches
        self.input_shape = input_shape

    def convert_bbx_to_feature_map(self, bbx, img_size):
        return convert_bbx_to_feature_map_batch([bbx], self.num_of_patches, img_size)[0]
