"""Shared helpers for the test-suite (synthetic inputs of SURVEY 8d, seeded weights)."""
import math
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def seeded_poolresnet_params(filters=64, seed=2, in_ch=3, num_blocks=10, stem_k=10, stem_s=8, head_k=6):
    """Default-initialised weights in the reference's construction order
    (models/PoolResnet.py:70-89: conv1, blocks[conv1, conv2], out) under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    p = {}
    c = torch.nn.Conv2d(in_ch, filters, stem_k, stride=stem_s, padding=stem_k - stem_s)
    p["conv1.weight"], p["conv1.bias"] = c.weight.detach(), c.bias.detach()
    for b in range(num_blocks):
        for n in ("conv1", "conv2"):
            c = torch.nn.Conv2d(filters, filters, 3, padding=1)
            p[f"residual_blocks.{b}.{n}.weight"] = c.weight.detach()
            p[f"residual_blocks.{b}.{n}.bias"] = c.bias.detach()
    c = torch.nn.Conv2d(filters, 5, head_k)
    p["out.weight"], p["out.bias"] = c.weight.detach(), c.bias.detach()
    return p


def synth_boxes(gen, kmin, kmax, size=480):
    """SURVEY 8d boxes: integer-valued (1,x,y,w,h) f32, log-uniform sizes clipped to the image."""
    k = int(torch.randint(kmin, kmax + 1, (1,), generator=gen))
    x = torch.randint(0, size, (k,), generator=gen).float()
    y = torch.randint(0, size, (k,), generator=gen).float()
    lw = torch.rand(k, generator=gen) * (math.log(240.0) - math.log(4.0)) + math.log(4.0)
    lh = torch.rand(k, generator=gen) * (math.log(240.0) - math.log(4.0)) + math.log(4.0)
    w = torch.minimum(torch.round(torch.exp(lw)), size - x).clamp(min=1)
    h = torch.minimum(torch.round(torch.exp(lh)), size - y).clamp(min=1)
    return torch.stack([torch.ones(k), x, y, w, h], dim=1)


def seeded_separable_params(filters=64, seed=6, in_ch=3, num_blocks=10, stem_k=10, stem_s=8, head_k=6):
    """Default-initialised weights in the reference's construction order (models/SeparableCNN.py:77-101: conv1,
    blocks[pointwise_conv1, depthwise_conv, pointwise_conv2], out) under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    p = {}
    torch.nn.Dropout2d(0.5)
    c = torch.nn.Conv2d(in_ch, filters, stem_k, stride=stem_s, padding=stem_k - stem_s)
    p["conv1.weight"], p["conv1.bias"] = c.weight.detach(), c.bias.detach()
    for b in range(num_blocks):
        pre = f"residual_blocks.{b}."
        p[pre + "pointwise_conv1.weight"] = torch.nn.Conv2d(filters, filters, 1, bias=False).weight.detach()
        p[pre + "depthwise_conv.weight"] = torch.nn.Conv2d(filters, filters, 3, padding=1, groups=filters,
                                                           bias=False).weight.detach()
        p[pre + "pointwise_conv2.weight"] = torch.nn.Conv2d(filters, filters, 1, bias=False).weight.detach()
    c = torch.nn.Conv2d(filters, 5, head_k)
    p["out.weight"], p["out.bias"] = c.weight.detach(), c.bias.detach()
    return p
