"""SSD -- mirror of the reference's ``models/SSD.py:14-255``: 3x3 stride-2 stem, nine + four ``SeparableResidualBlock``s
(two 3x3 convolutions, LeakyReLU 0.2, Dropout2d 0.25, 1x1 skip convolution where in != out, optional MaxPool2d), four
``Linear(C -> 5)`` heads over the 60 / 30 / 15 / 7 grids (4774 priors), sigmoid on the scores, ``apply_priors``.

Same constructor, ``forward(x, predict=torch.tensor(0))`` contract and ``state_dict`` keys (``input_normalizer.*``,
``feature_extractor.{k}.{pointwise_conv_skip,conv1,conv2}.*``, ``continue_layers.{i}.0.*``, ``extracting_layers.{i}.0.*``);
sub-modules are constructed in the reference's order, so ``torch.manual_seed`` gives the reference's initial weights.
The nn.Conv2d / nn.Linear sub-modules only HOLD the parameters; the arithmetic runs through ``SSDEngine`` (hand-written
sm_100a kernels on 64-channel planes).  There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..datasets.utils import ReduceSSDBoundingBoxes
from ..engine_ssd import SSDEngine
from .BaseSSDModel import BaseSSDModel


class SeparableResidualBlock(nn.Module):
    """Parameter holder for reference SSD.py:14-81."""

    def __init__(self, in_filters, out_filters, dropout=0.25, use_max_pool=False, bias=True):
        super().__init__()
        self.in_filters, self.out_filters = in_filters, out_filters
        self.filters_equal = in_filters == out_filters
        if not self.filters_equal:
            self.pointwise_conv_skip = nn.Conv2d(in_filters, out_filters, kernel_size=(1, 1), padding=0, bias=bias)
        self.conv1 = nn.Conv2d(in_filters, out_filters, kernel_size=(3, 3), padding=1, bias=bias)
        self.conv2 = nn.Conv2d(out_filters, out_filters, kernel_size=(3, 3), padding=1, bias=bias)
        self.use_max_pool = use_max_pool
        self.max_pool = nn.MaxPool2d(2)
        self.leaky_relu = nn.LeakyReLU(0.2)
        self.dropout2d = nn.Dropout2d(dropout)

    def forward(self, x):  # pragma: no cover - the engine runs the block
        raise RuntimeError("SeparableResidualBlock is executed by SSDEngine (CUDA only); call the parent model")


class _SSDFn(torch.autograd.Function):
    """Autograd bridge: forward / backward of the whole model as one node."""

    @staticmethod
    def forward(ctx, model, x, *params):
        eng = model.engine
        priors, mult = model._device_priors(x.device)
        with torch.cuda.device(x.device):
            pl = eng.forward(x, train=True, dropout=model.training, priors=priors, mult=mult)
        ctx.model, ctx.pl, ctx.generation = model, pl, pl["generation"]
        return pl["y"].clone()

    @staticmethod
    def backward(ctx, dy):
        model, pl = ctx.model, ctx.pl
        eng = model.engine
        if pl["generation"] != ctx.generation:
            raise RuntimeError("fd_b200 SSD: backward() after ANOTHER forward of the same batch size overwrote the saved "
                               "activations; call backward() before the next forward")
        priors, mult = model._device_priors(dy.device)
        with torch.cuda.device(dy.device):
            eng.run_backward(pl, dy.contiguous().float(), priors=priors, mult=mult)
        grads = [eng.grad_view(n).clone() for n in eng.param_names()]
        return (None, None, *grads)


class SSD(BaseSSDModel):
    def __init__(self, filters, input_shape, probability_threshold=0.5, iou_threshold=0.5, priors=None):
        super().__init__(filters, input_shape, probability_threshold=probability_threshold, iou_threshold=iou_threshold)
        self.patch_sizes = (60, 30, 15, 7)
        _, self.height, self.width = input_shape
        self.dropout2d = nn.Dropout2d(0.5)             # constructed by the reference, not used in forward (SSD.py:106,222-255)
        self.sigmoid = nn.Sigmoid()
        self.min_filters = filters
        self.max_filters = 16 * filters
        self.multiply_priors = torch.unsqueeze(
            torch.cat([torch.tensor(1 / ps).repeat(ps * ps) for ps in self.patch_sizes]), dim=1)      # SSD.py:111-116
        self.priors = priors if priors else self.calculate_priors()
        self.reduce_bounding_boxes = ReduceSSDBoundingBoxes(
            probability_threshold=probability_threshold, iou_threshold=iou_threshold, input_shape=self.input_shape,
            patch_sizes=self.patch_sizes, priors=self.priors)
        self.input_normalizer = nn.Conv2d(3, filters, kernel_size=(3, 3), stride=(2, 2), padding=1, bias=True)
        f = filters
        self.feature_extractor = nn.Sequential(
            SeparableResidualBlock(f, 2 * f, use_max_pool=True),
            SeparableResidualBlock(2 * f, 2 * f, use_max_pool=True),
            *[SeparableResidualBlock(2 * f, 2 * f, use_max_pool=False) for _ in range(6)],
            SeparableResidualBlock(2 * f, 4 * f, use_max_pool=False))
        continue_layers, extracting_layers = [], []
        for i, ps in enumerate(self.patch_sizes):                                                    # SSD.py:164-189
            in_filters = min(4 * f * (2 ** i), self.max_filters)
            out_filters = min(2 * in_filters, self.max_filters)
            continue_layers.append(nn.Sequential(SeparableResidualBlock(in_filters, out_filters, use_max_pool=i != 0)))
            extracting_layers.append(nn.Sequential(nn.Linear(in_features=out_filters, out_features=5)))
        self.continue_layers = nn.ModuleList(continue_layers)
        self.extracting_layers = nn.ModuleList(extracting_layers)
        self.avg_pooling = nn.AdaptiveAvgPool2d(2)
        self.engine = SSDEngine(filters, input_shape[0], input_shape[1], input_shape[2])
        if self.engine.patch_sizes != self.patch_sizes:
            raise NotImplementedError(f"input shape {input_shape} gives head grids {self.engine.patch_sizes}, the "
                                      f"reference's priors are laid out for {self.patch_sizes} (480 x 480 input)")
        self._dev_priors = None

    def calculate_priors(self):
        """SSD.py:192-204, verbatim arithmetic (the priors are data the kernels consume)."""
        priors_list = []
        for ps in self.patch_sizes:
            priors = torch.zeros((4, ps, ps))
            i, j = torch.where(priors[0] >= 0)
            priors[0, i, j] = priors[0, i, j] + 1 / ps * i
            priors[1, i, j] = priors[1, i, j] + 1 / ps * j
            priors = priors.permute(1, 2, 0).reshape(ps * ps, 4)
            priors_list.append(priors)
        return torch.cat(priors_list, dim=0)

    def _device_priors(self, device):
        if self._dev_priors is None or self._dev_priors[0].device != device:
            self._dev_priors = (self.priors.to(device).float().contiguous(),
                                self.multiply_priors.to(device).float().reshape(-1).contiguous())
        return self._dev_priors

    def _prep_input(self, x, predict):
        if predict:
            x = self._resize(x.to(next(self.parameters()).device) if not x.is_cuda else x)
            if x.dtype != torch.uint8:
                x = x / 255.0                                   # SSD.py:225 (uint8: fused into the stem kernel)
            if len(x.shape) == 3:
                x = torch.unsqueeze(x, 0)
        if not x.is_cuda:
            raise RuntimeError("fd_b200 models run on CUDA tensors only (no CPU fallback)")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        return x.contiguous()

    def forward(self, x: torch.Tensor, predict: torch.Tensor = torch.tensor(0)):
        is_predict = bool(predict == 1)
        x = self._prep_input(x, is_predict)
        params = dict(self.named_parameters())
        self.engine.bind(params)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params.values()):
            y = _SSDFn.apply(self, x, *[params[n] for n in self.engine.param_names()])
        else:
            priors, mult = self._device_priors(x.device)
            with torch.cuda.device(x.device):
                y = self.engine.forward(x, train=False, dropout=self.training, priors=priors, mult=mult)["y"].clone()
        if is_predict:
            return self.non_max_suppression(y)                   # SSD.py:253-254
        return y

    # ---- fused fast path: forward + ssd_loss (+ its gradient) + backward [+ all-reduce + Adam] (ModelMetaSSD.py:175)
    def train_step(self, x, y, neg_pos_ratio=10, optimizer=None, allreduce=None, num_pos_reduce=None):
        """Returns the loss (0-d device tensor); gradients land in ``p.grad`` (views of ``self.engine.gflat``)."""
        params = dict(self.named_parameters())
        self.engine.bind(params)
        x = self._prep_input(x, False)
        priors, mult = self._device_priors(x.device)
        pl = self.engine.train_step(x, y.float().contiguous(), priors, mult, neg_pos_ratio, dropout=self.training,
                                    allreduce=allreduce, optimizer=optimizer, num_pos_reduce=num_pos_reduce)
        for n, p in params.items():
            p.grad = self.engine.grad_view(n)
        return pl["loss"]

    def flat_optimizer(self, lr: float = 1e-4, capturable: bool = False):
        from ..optim import FlatAdam
        self.engine.bind(dict(self.named_parameters()))
        return FlatAdam(self.engine, lr=lr, capturable=capturable)


def bench_train_step(fd, dev, world, args, barrier, par, timed_graph_region, capture, synth_batch):
    """bench.py leg `train_ssd` (BASELINE config 5): SSD(filters=16) forward + ssd_loss(.., 10) + backward + Adam, 16
    images per GPU (128 over 8 GPUs), targets encoded at the four scales from synthetic boxes (< 120 per image)."""
    import torch.distributed as dist
    Bs = 16
    torch.manual_seed(2)
    m = SSD(filters=16, input_shape=(3, 480, 480)).to(dev).train()
    m.engine.bind(dict(m.named_parameters()))
    par.broadcast_flat(m.engine.pflat)
    rank = dist.get_rank() if dist.is_initialized() else 0
    x_cpu, boxes = synth_batch(Bs, seed_img=10 + 2 * rank, seed_box=11 + 2 * rank, kmin=1, kmax=119)
    gt = fd.datasets.WIDERFace.dataset_ssd.convert_bbx_to_feature_maps_batch(boxes, (480, 480), device=dev)
    x = x_cpu.to(dev)
    opt = m.flat_optimizer(lr=1e-4, capturable=True)
    opt._ensure_state()
    priors, mult = m._device_priors(dev)
    ar = par.allreduce_grads if world > 1 else None

    def npos_reduce(n):
        if world > 1:
            dist.all_reduce(n)
        return n

    def step():
        return m.engine.train_step(x, gt, priors, mult, 10, dropout=True, allreduce=ar, optimizer=opt,
                                   num_pos_reduce=npos_reduce if world > 1 else None)

    n0 = fd.native.launch_count()
    step()
    launches = fd.native.launch_count() - n0
    if world == 1:
        g, pl, launches = capture(step)
        run = g.replay
    else:
        run = step                       # the NCCL exchanges (num_pos + gradients) stay outside a graph
    ms, _, _ = timed_graph_region(run, max(3, min(args.steps, 10)), 2, barrier, par, dev)
    flops_img = 3 * 14.4e9
    return {"metric": "train_images_per_sec", "value": world * Bs / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
            "launches_per_step": launches, "cuda_graph": world == 1,
            "achieved_tflops_algorithmic": flops_img * Bs / (ms * 1e-3) / 1e12,
            "workload": "SSD(filters=16, 480x480; BASELINE config 5) train step: forward + ssd_loss(neg_pos_ratio 10) + "
                        "backward" + (" + num_pos and gradient all-reduce (NCCL)" if world > 1 else "") + " + Adam, 16 "
                        "images per GPU; 64-channel planes (channel counts 16/32 zero-padded), the 128 / 256-channel blocks on the "
                        "cta_group::2 wide kernels"}
