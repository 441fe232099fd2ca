"""PoolResnet -- mirror of the reference's ``models/PoolResnet.py:11-105``.

Same constructor, same ``forward(x, predict=torch.tensor(0))`` contract and the same ``state_dict``
keys (``conv1.*``, ``residual_blocks.{k}.conv{1,2}.*``, ``out.*``), so reference checkpoints load with
``strict=True``.  The nn.Conv2d sub-modules only *hold* the parameters; the arithmetic runs through
``BackboneEngine`` (hand-written sm_100a kernels).  There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..engine import BackboneEngine
from ..engine_planar import PlanarEngine
from .BaseModel import BaseModel


class ResidualBlock(nn.Module):
    """Parameter holder for reference PoolResnet.py:11-43 (two 3x3 convs, LeakyReLU 0.2, Dropout2d, skip,
    conditional MaxPool2d(2))."""

    def __init__(self, filters, num_of_patches, dropout=0.25):
        super().__init__()
        self.num_of_patches = num_of_patches
        self.conv1 = nn.Conv2d(filters, filters, kernel_size=(3, 3), padding=1)
        self.conv2 = nn.Conv2d(filters, filters, kernel_size=(3, 3), padding=1)
        self.max_pool = nn.MaxPool2d(2)
        self.leaky_relu = nn.LeakyReLU(0.2)
        self.dropout2d = nn.Dropout2d(dropout)

    def forward(self, x):  # pragma: no cover - the engine runs the block
        raise RuntimeError("ResidualBlock is executed by BackboneEngine (CUDA only); call the parent model")


class _BackboneFn(torch.autograd.Function):
    """Autograd bridge: forward/backward of the whole backbone as one node."""

    @staticmethod
    def forward(ctx, model, x, *params):
        eng = model.engine
        pl = eng.forward(x, train=True, dropout=model.training)
        ctx.eng, ctx.pl = eng, pl
        return pl.y.clone()

    @staticmethod
    def backward(ctx, dy):
        eng, pl = ctx.eng, ctx.pl
        eng.run_backward(pl, dy.contiguous().float())
        grads = [eng.grad_view(n).clone() for n in eng.param_names()]
        return (None, None, *grads)


class GridBackbone(BaseModel):
    """Shared implementation of PoolResnet / Resnet: stem conv -> residual blocks -> head conv -> sigmoid."""

    def _build(self, filters, input_shape, num_of_residual_blocks, stem_k, stem_s, stem_pad, head_k, head_pad,
               pool_rule, block_patches):
        self.dropout2d = nn.Dropout2d(0.5)
        self.conv1 = nn.Conv2d(input_shape[0], filters, kernel_size=(stem_k, stem_k), stride=(stem_s, stem_s),
                               padding=stem_pad)
        self.residual_blocks = nn.Sequential(
            *[ResidualBlock(filters=filters, num_of_patches=block_patches) for _ in range(num_of_residual_blocks)])
        self.out = nn.Conv2d(filters, 5, stride=(1, 1), kernel_size=(head_k, head_k), padding=head_pad)
        self.sigmoid = nn.Sigmoid()
        # 64 channels: the fused tensor-core engine; 128, 192, ...: the same kernels on 64-channel planes (engine_planar)
        Engine = BackboneEngine if filters == 64 else PlanarEngine
        self.engine = Engine(filters, input_shape[0], input_shape[1], input_shape[2], num_of_residual_blocks,
                             stem_k, stem_s, stem_pad, head_k, head_pad, pool_rule)

    def _prep_input(self, x: torch.Tensor, predict: bool) -> torch.Tensor:
        if predict:
            x = self._resize(x)                       # PoolResnet.py:94-95
            if x.dtype != torch.uint8:
                x = x / 255.0
            if len(x.shape) == 3:
                x = torch.unsqueeze(x, 0)
        if not x.is_cuda:
            raise RuntimeError("fd_b200 models run on CUDA tensors only (no CPU fallback)")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        return x.contiguous()

    def forward(self, x: torch.Tensor, predict: torch.Tensor = torch.tensor(0)):
        is_predict = bool(predict == 1)
        x = self._prep_input(x, is_predict)           # uint8 input: the /255 is fused into the stem kernel
        params = dict(self.named_parameters())
        self.engine.bind(params)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params.values()):
            plist = [params[n] for n in self.engine.param_names()]
            y = _BackboneFn.apply(self, x, *plist)
        else:
            y = self.engine.forward(x, train=False, dropout=self.training).y.clone()
        if is_predict:
            return self.single_non_max_suppression(y[0])   # PoolResnet.py:103-104
        return y

    # ---- fused fast path (one call = forward + summed YoloLoss + backward), see engine.train_step
    def train_step(self, x: torch.Tensor, gt: torch.Tensor, optimizer=None, allreduce=None):
        """Returns the summed loss (0-d device tensor); gradients land in ``p.grad`` (views of the flat
        gradient buffer ``self.engine.gflat``).  ``allreduce`` (parallel.PeerAllReduce / parallel.allreduce_grads)
        sums the gradient over the data-parallel ranks; ``optimizer`` (``self.flat_optimizer()``) then applies the
        reference's Adam update (models/ModelMeta.py:104-112) in the same call."""
        params = dict(self.named_parameters())
        self.engine.bind(params)
        pl = self.engine.train_step(self._prep_input(x, False), gt.float().contiguous(), dropout=self.training,
                                    allreduce=allreduce, optimizer=optimizer)
        for n, p in params.items():
            p.grad = self.engine.grad_view(n)
        return pl.loss.sum()

    def flat_optimizer(self, lr: float = 1e-4, capturable: bool = False):
        """One-kernel Adam over the flat parameter buffer (optim.FlatAdam); lr default = ModelMeta's (ModelMeta.py:86)."""
        from ..optim import FlatAdam
        if not isinstance(self.engine, BackboneEngine):
            raise NotImplementedError("flat_optimizer needs the flat parameter buffer of the 64-channel engine; use "
                                      "torch.optim.Adam(model.parameters()) (ModelMeta.configure_optimizers) for wider models")
        self.engine.bind(dict(self.named_parameters()))
        return FlatAdam(self.engine, lr=lr, capturable=capturable)


class PoolResnet(GridBackbone):
    def __init__(self, filters, input_shape, num_of_patches, num_of_residual_blocks=10, probability_threshold=0.5,
                 iou_threshold=0.5, pretrained=False, input_kernel_size=10, input_stride=8, output_kernel_size=6,
                 output_padding=0):
        super().__init__(filters, input_shape, num_of_patches=num_of_patches,
                         probability_threshold=probability_threshold, iou_threshold=iou_threshold)
        self.pretrained = pretrained
        S = self.num_of_patches
        self._build(filters, input_shape, num_of_residual_blocks, input_kernel_size, input_stride,
                    input_kernel_size - input_stride,            # PoolResnet.py:75
                    output_kernel_size, output_padding,
                    pool_rule=lambda h: h > 2 * S,               # PoolResnet.py:41
                    block_patches=S)
