"""PoolResnet -- mirror of the reference's ``models/PoolResnet.py:11-105``.

Same constructor, same ``forward(x, predict=torch.tensor(0))`` contract and the same ``state_dict``
keys (``conv1.*``, ``residual_blocks.{k}.conv{1,2}.*``, ``out.*``), so reference checkpoints load with
``strict=True``.  The nn.Conv2d sub-modules only *hold* the parameters; the arithmetic runs through
``BackboneEngine`` (hand-written sm_100a kernels).  There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..engine import BackboneEngine, PaddedBackboneEngine
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
        ctx.eng, ctx.pl, ctx.generation = eng, pl, pl.generation
        return pl.y.clone()

    @staticmethod
    def backward(ctx, dy):
        eng, pl = ctx.eng, ctx.pl
        if pl.generation != ctx.generation:
            # the saved activations / masks live in the engine's per-batch-size plan, not in the autograd graph
            raise RuntimeError(
                "fd_b200 backbone: backward() after ANOTHER forward of the same batch size overwrote the saved "
                "activations (gradient accumulation over micro-batches / SAM closures: call backward() before the "
                "next forward, or run the intermediate forward under torch.no_grad())")
        eng.run_backward(pl, dy.contiguous().float())
        grads = [eng.grad_view(n).clone() for n in eng.param_names()]
        return (None, None, *grads)


class GraphedTrainStep:
    """See ``GridBackbone.graphed_train_step``."""

    def __init__(self, model, B, image_dtype, optimizer, allreduce, n_buffers):
        eng = model.engine
        dev = next(model.parameters()).device
        C, H, W = model.input_shape
        if optimizer is not None and not getattr(optimizer, "capturable", False):
            raise ValueError("an optimizer inside a CUDA graph must keep its step count on the device: "
                             "use model.flat_optimizer(capturable=True)")
        with torch.cuda.device(dev):
            self.x = [torch.zeros((B, C, H, W), dtype=image_dtype, device=dev) for _ in range(n_buffers)]
            self.gt = [torch.zeros((B, 5, eng.So_h, eng.So_w), dtype=torch.float32, device=dev) for _ in range(n_buffers)]
            if optimizer is not None:
                optimizer._ensure_state()
            dropout = model.training
            # warm-up (plan allocation, tensor-map encoding) WITHOUT the optimizer / exchange: the weights are untouched
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                eng.train_step(self.x[0], self.gt[0], dropout=dropout)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            self.graphs, self.plan = [], None
            for i in range(n_buffers):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.plan = eng.train_step(self.x[i], self.gt[i], dropout=dropout, allreduce=allreduce,
                                               optimizer=optimizer)
                self.graphs.append(g)

    def replay(self, i: int = 0) -> torch.Tensor:
        self.graphs[i].replay()
        return self.plan.loss


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
        # fewer than 64 (the 'small' checkpoint is 32): the 64-channel engine on zero-padded weights
        Engine = BackboneEngine if filters == 64 else PaddedBackboneEngine if filters < 64 else PlanarEngine
        self.engine = Engine(filters, input_shape[0], input_shape[1], input_shape[2], num_of_residual_blocks,
                             stem_k, stem_s, stem_pad, head_k, head_pad, pool_rule)

    def _prep_input(self, x: torch.Tensor, predict: bool) -> torch.Tensor:
        if predict:
            x = self._resize(self._to_model_device(x))   # PoolResnet.py:94-95
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

    def graphed_train_step(self, batch_size: int, image_dtype=torch.float32, optimizer=None, allreduce=None,
                           n_buffers: int = 2):
        """The train step (forward + summed YoloLoss + backward [+ all-reduce] [+ Adam]) captured ONCE per static input
        buffer into a CUDA graph: ``step.x[i]`` / ``step.gt[i]`` are device buffers the caller fills (e.g. with
        ``copy_(pinned_host_batch, non_blocking=True)`` on a copy stream), ``step.replay(i)`` launches the whole step
        as one graph and returns the per-image losses ``[B]`` (device tensor, valid until the next replay).  With
        ``n_buffers = 2`` the copy of batch k+1 overlaps the compute of batch k.  ``optimizer`` must be
        ``flat_optimizer(capturable=True)``; gradients land in ``p.grad`` (views of the flat gradient buffer)."""
        params = dict(self.named_parameters())
        self.engine.bind(params)
        step = GraphedTrainStep(self, batch_size, image_dtype, optimizer, allreduce, n_buffers)
        for n, p in params.items():
            p.grad = self.engine.grad_view(n)
        return step

    def flat_optimizer(self, lr: float = 1e-4, capturable: bool = False):
        """One-kernel Adam over the flat parameter buffer (optim.FlatAdam); lr default = ModelMeta's (ModelMeta.py:86)."""
        from ..optim import FlatAdam
        if not hasattr(self.engine, "opt_params"):
            raise NotImplementedError("flat_optimizer needs an engine with a flat parameter buffer; use "
                                      "torch.optim.Adam(model.parameters()) (ModelMeta.configure_optimizers)")
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
