"""CPU oracle for the convolutional backbone -- TEST INFRASTRUCTURE ONLY.

A torch-fp32 *functional* restatement of the reference's PoolResnet / Resnet /
SeparableCNN forward passes, of the training-step definition (forward, sum of
per-image ``yolo_loss``, backward) and of the optimizer update (``adam_update``).  Floating-point kernels keep a torch fp32 reference;
this is it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline legs may import this module.

Parity status: PINNED against the real reference modules run in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/backbone_*.npz``,
checked by ``tests/test_oracle_golden.py``): logits identical (max-abs-diff 0.0,
same ATen kernels), loss and gradients identical; likewise Resnet
(``make_golden_extra.py``), SeparableCNN (``make_golden_sep.py``) and the
128-channel PoolResnet / SeparableCNN (``make_golden_wide.py``).

Reference citations are ``file:line`` under ``/root/reference``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


def residual_block(x, w1, b1, w2, b2, pool: bool, drop_scale: Optional[torch.Tensor] = None):
    """models/PoolResnet.py:33-43 (same body in models/Resnet.py:30-40).

    conv3x3 p1 -> LeakyReLU(0.2) -> conv3x3 p1 -> LeakyReLU(0.2) -> Dropout2d -> + skip ->
    optional MaxPool2d(2).  ``drop_scale`` is an explicit ``[B,C,1,1]`` Dropout2d multiplier
    (0 or 1/(1-p)); ``None`` = eval mode.
    """
    y = F.leaky_relu(F.conv2d(x, w1, b1, padding=1), 0.2)
    y = F.leaky_relu(F.conv2d(y, w2, b2, padding=1), 0.2)
    if drop_scale is not None:
        y = y * drop_scale
    y = y + x
    if pool:
        y = F.max_pool2d(y, 2)
    return y


def poolresnet_forward(x: torch.Tensor, p: Params, num_of_patches: int, num_blocks: int = 10,
                       input_stride: int = 8, input_kernel_size: int = 10, output_padding: int = 0,
                       drop_scales: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    """models/PoolResnet.py:93-105 with ``predict == 0``.

    stem padding = kernel - stride (PoolResnet.py:75); a block pools iff its (pre-pool)
    height is > 2*S (PoolResnet.py:41); head = valid conv + sigmoid (PoolResnet.py:83-89,101-102).
    ``drop_scales``: optional list of ``num_blocks + 1`` multipliers (blocks, then the
    Dropout2d(0.5) before the head, PoolResnet.py:100).
    """
    y = F.conv2d(x, p["conv1.weight"], p["conv1.bias"], stride=input_stride,
                 padding=input_kernel_size - input_stride)
    for b in range(num_blocks):
        pre = f"residual_blocks.{b}."
        ds = None if drop_scales is None else drop_scales[b]
        y = residual_block(y, p[pre + "conv1.weight"], p[pre + "conv1.bias"],
                           p[pre + "conv2.weight"], p[pre + "conv2.bias"],
                           pool=y.shape[2] > 2 * num_of_patches, drop_scale=ds)
    if drop_scales is not None and drop_scales[num_blocks] is not None:
        y = y * drop_scales[num_blocks]
    y = F.conv2d(y, p["out.weight"], p["out.bias"], padding=output_padding)
    return torch.sigmoid(y)


def resnet_forward(x: torch.Tensor, p: Params, num_of_patches: int, num_blocks: int = 10,
                   drop_scales: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    """models/Resnet.py:89-99: 3x3 s2 p1 stem, blocks pool while H > S (Resnet.py:38), Dropout2d(0.5) (:95),
    3x3 p1 head + sigmoid.  ``drop_scales`` as in ``poolresnet_forward`` (None = eval mode)."""
    y = F.conv2d(x, p["conv1.weight"], p["conv1.bias"], stride=2, padding=1)
    for b in range(num_blocks):
        pre = f"residual_blocks.{b}."
        ds = None if drop_scales is None else drop_scales[b]
        y = residual_block(y, p[pre + "conv1.weight"], p[pre + "conv1.bias"],
                           p[pre + "conv2.weight"], p[pre + "conv2.bias"],
                           pool=y.shape[2] > num_of_patches, drop_scale=ds)
    if drop_scales is not None and drop_scales[num_blocks] is not None:
        y = y * drop_scales[num_blocks]
    y = F.conv2d(y, p["out.weight"], p["out.bias"], padding=1)
    return torch.sigmoid(y)


def resize_bilinear(x: torch.Tensor, size) -> torch.Tensor:
    """``transforms.Resize(size)`` of models/PoolResnet.py:91,95 / models/BaseModel.py:64 under the reference's pinned
    torchvision 0.11.2 (and inside the official TorchScript archives, whose antialias branch is dead): bilinear,
    align_corners=False, no antialias; uint8 goes through float32, ``round`` and back (functional_tensor.py resize ->
    _cast_squeeze_in / _cast_squeeze_out)."""
    squeeze = x.dim() == 3
    xb = x.unsqueeze(0) if squeeze else x
    if tuple(xb.shape[-2:]) == tuple(size):
        return x
    y = F.interpolate(xb.float(), size=tuple(size), mode="bilinear", align_corners=False)
    if x.dtype == torch.uint8:
        y = torch.round(y).to(torch.uint8)
    return y[0] if squeeze else y


def yolo_loss_torch(pred_fm: torch.Tensor, gt_fm: torch.Tensor) -> torch.Tensor:
    """losses/YoloLoss.py:4-44 restated for autograd (same op order, same swapped x/y)."""
    S = pred_fm.shape[1]
    p = pred_fm.reshape(5, -1)
    if torch.nansum(p):                                   # YoloLoss.py:8-9
        p = torch.nan_to_num(p, nan=0.1)
    g = gt_fm.reshape(5, -1)
    g0 = g[0]
    xy = 3 * g0 * ((g[1] - p[2]) ** 2 + (g[2] - p[1]) ** 2)      # YoloLoss.py:17-18,27-29
    wh = 3 * g0 * ((g[3] ** 0.5 - p[3] ** 0.5) ** 2 + (g[4] ** 0.5 - p[4] ** 0.5) ** 2)
    cf = (g0 + (1 - g0) * (1 / S)) * (g0 - p[0]) ** 2            # YoloLoss.py:25,36-38
    return torch.sum(xy + wh + cf)


def train_step(x: torch.Tensor, y: torch.Tensor, p: Params, num_of_patches: int,
               forward=poolresnet_forward, **fw):
    """The train-step definition (models/ModelMeta.py:141,173-176): ``y_hat = model(x)``;
    ``loss = sum_i yolo_loss(y_hat[i], y[i])`` (a SUM, ModelMeta.py:215 has the mean commented
    out); ``loss.backward()``.  Returns (y_hat, loss, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    y_hat = forward(x, leaves, num_of_patches, **fw)
    loss = 0
    for i in range(y.shape[0]):                           # per-image python loop, as the reference
        loss = loss + yolo_loss_torch(y_hat[i], y[i])
    loss.backward()
    grads = {k: v.grad.detach() for k, v in leaves.items()}
    return y_hat.detach(), loss.detach(), grads


def adam_update(p: Params, grads: Params, state: dict, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8):
    """The reference's optimizer step (models/ModelMeta.py:104-112, 12-82): ``SAMSGD`` perturbs the weights by
    ``rho * g / |g|`` and immediately removes the perturbation again WITHOUT re-evaluating the closure, so the update
    that remains is its base ``torch.optim._multi_tensor.Adam(lr)``.  Restated with torch.optim.Adam
    on the same tensors; ``state`` carries the optimizer between calls.  Updates ``p`` in place."""
    if "opt" not in state:
        state["leaves"] = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        state["opt"] = torch.optim.Adam(list(state["leaves"].values()), lr=lr, betas=betas, eps=eps)
    for k, leaf in state["leaves"].items():
        leaf.grad = grads[k].detach().clone()
    state["opt"].step()
    for k, leaf in state["leaves"].items():
        p[k] = leaf.detach().clone()
    return p


def separable_block(x: torch.Tensor, w_pw1: torch.Tensor, w_dw: torch.Tensor, w_pw2: torch.Tensor, pool: bool,
                    slope: float = 0.2, drop_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """models/SeparableCNN.py:40-51 (eval: Dropout2d is the identity): 1x1 -> lrelu -> depthwise 3x3 pad 1 -> lrelu ->
    1x1 -> + skip -> MaxPool2d(2) while H > num_of_patches.  All three convolutions have bias=False (:11,19,27,35)."""
    y = F.leaky_relu(F.conv2d(x, w_pw1), slope)
    y = F.leaky_relu(F.conv2d(y, w_dw, padding=1, groups=w_dw.shape[0]), slope)
    y = F.conv2d(y, w_pw2)
    if drop_scale is not None:                     # Dropout2d on the pw2 output, before the skip add (:47-48)
        y = y * drop_scale
    y = y + x
    return F.max_pool2d(y, 2) if pool else y


def separable_forward(x: torch.Tensor, p: Params, num_of_patches: int = 16, num_blocks: int = 10, block_patches: int = 16,
                      drop_scales: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    """models/SeparableCNN.py:104-117 (eval, predict == 0): 10x10 stride-8 pad-2 stem, blocks that pool while
    H > 16 (the constructor hard-wires num_of_patches=16, :71,87-91), 6x6 valid head, sigmoid."""
    k = p["conv1.weight"].shape[2]
    y = F.conv2d(x, p["conv1.weight"], p["conv1.bias"], stride=8, padding=k - 8)
    for b in range(num_blocks):
        pre = f"residual_blocks.{b}."
        y = separable_block(y, p[pre + "pointwise_conv1.weight"], p[pre + "depthwise_conv.weight"],
                            p[pre + "pointwise_conv2.weight"], pool=y.shape[2] > block_patches,
                            drop_scale=None if drop_scales is None else drop_scales[b])
    if drop_scales is not None and drop_scales[num_blocks] is not None:
        y = y * drop_scales[num_blocks]                                     # Dropout2d(0.5) before the head (:109)
    y = F.conv2d(y, p["out.weight"], p["out.bias"])
    return torch.sigmoid(y)


# ------------------------------------------------------------------------------------------------ MobilenetV3 backbone
# timm 0.5.4 `tf_mobilenetv3_small_100` minus its last five children (models/MobilenetV3Backbone.py:33-39) + the
# Conv2d(576 -> 5, 3x3, pad 1) head + sigmoid (:40-46,57-58).  timm is not installed; the graph below restates the
# code stored in the official TorchScript archive (saved_models/official/MobilenetV3Backbone/medium_model_15x15_480.pth:
# code/__torch__/timm/models/efficientnet_blocks.py, .../layers/conv2d_same.py, .../layers/padding.py).
# (kind, kernel, stride, expanded channels, out channels, activation, SE reduced channels or 0)
MBV3_STAGES = (
    (("ds", 3, 2, 16, 16, "relu", 8),),
    (("ir", 3, 2, 72, 24, "relu", 0), ("ir", 3, 1, 88, 24, "relu", 0)),
    (("ir", 5, 2, 96, 40, "hswish", 24), ("ir", 5, 1, 240, 40, "hswish", 64), ("ir", 5, 1, 240, 40, "hswish", 64)),
    (("ir", 5, 1, 120, 48, "hswish", 32), ("ir", 5, 1, 144, 48, "hswish", 40)),
    (("ir", 5, 2, 288, 96, "hswish", 72), ("ir", 5, 1, 576, 96, "hswish", 144), ("ir", 5, 1, 576, 96, "hswish", 144)),
    (("cba", 1, 1, 576, 576, "hswish", 0),),
)
MBV3_BN_EPS = 1e-3


def _same_pad(size: int, k: int, s: int):
    """timm layers/padding.py get_same_padding + pad_same: TF 'SAME' -- total = max((ceil(i/s)-1)*s + k - i, 0), the
    extra pixel goes to the bottom / right."""
    total = max((-(-size // s) - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _conv_same(x, w, stride, groups=1):
    """Conv2dSame.forward (stride-2 layers of the tf_ variant); stride-1 layers are plain Conv2d(padding=k//2)."""
    k = w.shape[-1]
    if stride == 1:
        return F.conv2d(x, w, None, 1, k // 2, 1, groups)
    pt, pb = _same_pad(x.shape[-2], k, stride)
    pl, pr = _same_pad(x.shape[-1], k, stride)
    return F.conv2d(F.pad(x, [pl, pr, pt, pb]), w, None, stride, 0, 1, groups)


def _bn(x, p: Params, pre: str):
    return F.batch_norm(x, p[pre + "running_mean"], p[pre + "running_var"], p[pre + "weight"], p[pre + "bias"],
                        False, 0.0, MBV3_BN_EPS)


def _act(x, name):
    return F.relu(x) if name == "relu" else F.hardswish(x) if name == "hswish" else x


def _se(x, p: Params, pre: str):
    """efficientnet_blocks.py SqueezeExcite: mean over (H, W) -> conv_reduce -> ReLU -> conv_expand -> Hardsigmoid gate."""
    s = x.mean((2, 3), keepdim=True)
    s = F.relu(F.conv2d(s, p[pre + "conv_reduce.weight"], p[pre + "conv_reduce.bias"]))
    s = F.conv2d(s, p[pre + "conv_expand.weight"], p[pre + "conv_expand.bias"])
    return x * F.hardsigmoid(s)


def mobilenetv3_features(x: torch.Tensor, p: Params) -> torch.Tensor:
    """feature_extractor of models/MobilenetV3Backbone.py:33-39 in eval mode -> [B,576,H/32,W/32]."""
    fe = "feature_extractor."
    y = _act(_bn(_conv_same(x, p[fe + "0.weight"], 2), p, fe + "1."), "hswish")
    for si, stage in enumerate(MBV3_STAGES):
        for bi, (kind, k, s, cexp, cout, act, se) in enumerate(stage):
            pre = f"{fe}3.{si}.{bi}."
            inp = y
            if kind == "cba":
                y = _act(_bn(F.conv2d(y, p[pre + "conv.weight"]), p, pre + "bn1."), act)
                continue
            if kind == "ds":        # DepthwiseSeparableConv: dw -> bn -> act -> se -> pw -> bn (no act), skip if same shape
                y = _act(_bn(_conv_same(y, p[pre + "conv_dw.weight"], s, groups=y.shape[1]), p, pre + "bn1."), act)
                if se:
                    y = _se(y, p, pre + "se.")
                y = _bn(F.conv2d(y, p[pre + "conv_pw.weight"]), p, pre + "bn2.")
            else:                   # InvertedResidual: pw -> bn -> act -> dw -> bn -> act -> se -> pwl -> bn, skip
                y = _act(_bn(F.conv2d(y, p[pre + "conv_pw.weight"]), p, pre + "bn1."), act)
                y = _act(_bn(_conv_same(y, p[pre + "conv_dw.weight"], s, groups=y.shape[1]), p, pre + "bn2."), act)
                if se:
                    y = _se(y, p, pre + "se.")
                y = _bn(F.conv2d(y, p[pre + "conv_pwl.weight"]), p, pre + "bn3.")
            if s == 1 and inp.shape[1] == y.shape[1]:
                y = y + inp
    return y


def mobilenetv3_forward(x: torch.Tensor, p: Params) -> torch.Tensor:
    """models/MobilenetV3Backbone.py:50-60 with predict == 0, eval mode: features -> 3x3 pad-1 head -> sigmoid."""
    y = mobilenetv3_features(x, p)
    return torch.sigmoid(F.conv2d(y, p["out.weight"], p["out.bias"], padding=1))


# ------------------------------------------------------------------------------------------------ SSD model
SSD_PATCH_SIZES = (60, 30, 15, 7)


def ssd_block(x, p: Params, pre: str, pool: bool, drop_scale: Optional[torch.Tensor] = None):
    """models/SSD.py:63-81 SeparableResidualBlock.forward: optional 1x1 skip conv, conv3x3 -> lrelu -> conv3x3 -> lrelu ->
    Dropout2d (explicit [B,C,1,1] multiplier or None = eval) -> + skip -> optional MaxPool2d(2)."""
    skip = x
    if (pre + "pointwise_conv_skip.weight") in p:
        skip = F.conv2d(x, p[pre + "pointwise_conv_skip.weight"], p[pre + "pointwise_conv_skip.bias"])
    y = F.leaky_relu(F.conv2d(x, p[pre + "conv1.weight"], p[pre + "conv1.bias"], padding=1), 0.2)
    y = F.leaky_relu(F.conv2d(y, p[pre + "conv2.weight"], p[pre + "conv2.bias"], padding=1), 0.2)
    if drop_scale is not None:
        y = y * drop_scale
    y = y + skip
    return F.max_pool2d(y, 2) if pool else y


def ssd_priors_torch():
    """models/SSD.py:111-116,192-204: (multiply_priors [P,1], priors [P,4])."""
    mult = torch.unsqueeze(torch.cat([torch.tensor(1 / ps).repeat(ps * ps) for ps in SSD_PATCH_SIZES]), dim=1)
    pri = []
    for ps in SSD_PATCH_SIZES:
        t = torch.zeros((4, ps, ps))
        i, j = torch.where(t[0] >= 0)
        t[0, i, j] = t[0, i, j] + 1 / ps * i
        t[1, i, j] = t[1, i, j] + 1 / ps * j
        pri.append(t.permute(1, 2, 0).reshape(ps * ps, 4))
    return mult, torch.cat(pri, dim=0)


def ssd_forward(x: torch.Tensor, p: Params, drop_scales: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    """models/SSD.py:222-255 with predict == 0 -> [B,4774,5] rows (sigmoid score, x/ps + i/ps, y/ps + j/ps, w, h).
    ``drop_scales``: 13 Dropout2d multipliers (9 feature-extractor blocks, then the 4 continue blocks) or None."""
    bs = x.shape[0]
    y = F.conv2d(x, p["input_normalizer.weight"], p["input_normalizer.bias"], stride=2, padding=1)
    k = 0
    for b in range(9):
        y = ssd_block(y, p, f"feature_extractor.{b}.", pool=b < 2, drop_scale=None if drop_scales is None else drop_scales[k])
        k += 1
    scores, bbxs = [], []
    for i in range(4):
        y = ssd_block(y, p, f"continue_layers.{i}.0.", pool=i != 0, drop_scale=None if drop_scales is None else drop_scales[k])
        k += 1
        z = F.linear(y.permute(0, 2, 3, 1).contiguous(), p[f"extracting_layers.{i}.0.weight"], p[f"extracting_layers.{i}.0.bias"])
        z = z.reshape(bs, -1, 5)
        scores.append(z[..., :1])
        bbxs.append(z[..., 1:5])
    out = torch.cat([torch.sigmoid(torch.cat(scores, dim=1)), torch.cat(bbxs, dim=1)], dim=2)
    mult, pri = ssd_priors_torch()
    s = torch.clone(out).float()                                        # apply_priors, SSD.py:206-220
    s[..., 1:2] = s[..., 1:2] * mult.repeat(repeats=(bs, 1, 1))
    s[..., 2:3] = s[..., 2:3] * mult.repeat(repeats=(bs, 1, 1))
    s[..., 1:5] = s[..., 1:5] + pri.repeat(repeats=(bs, 1, 1))
    return s


def ssd_loss_torch(confidence, predicted_locations, labels, gt_locations, neg_pos_ratio):
    """losses/SSDLoss.py:27-86 restated for autograd (hard-negative mining, clamped BCE, smooth L1, / num_pos)."""
    import math
    with torch.no_grad():
        loss = -torch.log(confidence)
        pos_mask = labels > 0
        num_neg = pos_mask.long().sum(dim=1, keepdim=True) * neg_pos_ratio
        loss[pos_mask] = -math.inf
        _, indexes = loss.sort(dim=1, descending=True, stable=True)
        _, orders = indexes.sort(dim=1)
        mask = pos_mask | (orders < num_neg)
    c = confidence[mask].clamp(1e-7, 1 - 1e-7)
    t = torch.round(labels[mask])
    cls = torch.sum(-1 * (t * torch.log(c) + (1 - t) * torch.log(1 - c)))
    pl = predicted_locations[pos_mask, :].reshape(-1, 4)
    gl = gt_locations[pos_mask, :].reshape(-1, 4)
    return (F.smooth_l1_loss(pl, gl, reduction="sum") + cls) / gl.size(0)


def ssd_train_step(x, y, p: Params, neg_pos_ratio: int = 10, drop_scales=None):
    """models/ModelMetaSSD.py:143,175: y_hat = model(x); loss = ssd_loss(y_hat[:,:,0], y_hat[:,:,1:], y[:,:,0], y[:,:,1:], 10);
    loss.backward().  Returns (y_hat, loss, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    y_hat = ssd_forward(x, leaves, drop_scales)
    loss = ssd_loss_torch(y_hat[:, :, 0], y_hat[:, :, 1:], y[:, :, 0], y[:, :, 1:], neg_pos_ratio)
    loss.backward()
    return y_hat.detach(), loss.detach(), {k: v.grad.detach() for k, v in leaves.items()}
