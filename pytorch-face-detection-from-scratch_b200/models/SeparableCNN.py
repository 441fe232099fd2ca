"""SeparableCNN -- mirror of the reference's ``models/SeparableCNN.py:10-117`` (depthwise-separable backbone).

Same constructor, ``forward(x, predict=torch.tensor(0))`` contract and ``state_dict`` keys (``conv1.*``,
``residual_blocks.{k}.pointwise_conv1 / depthwise_conv / pointwise_conv2.weight``, ``out.*``).  Like the reference,
``num_of_patches`` is hard-wired to 16 (SeparableCNN.py:71) although the head emits a 10x10 map: the blocks pool while
H > 16 and the decoder works with 30-pixel patches -- reproduced, not fixed.

The forward (BASELINE config 4: inference + NMS) runs on hand-written sm_100a kernels: the stem and head kernels of
the residual backbones, and ONE fused kernel per separable block (``fd_sepblock_fwd``: 1x1 -> LeakyReLU -> depthwise 3x3 ->
LeakyReLU -> 1x1 -> + skip -> MaxPool2d(2) where the block pools; intermediates in shared memory).  Inference only:
the backward of this backbone is not built (``train()`` mode raises).  ``filters == 64`` runs the fused block kernel,
``filters == 128`` (the reference's ``__main__``) the 64-channel kernels on two channel planes.  There is no CPU path.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .. import ops
from .BaseModel import BaseModel

BF16, F32 = torch.bfloat16, torch.float32


class ResidualBlock(nn.Module):
    """Parameter holder for reference SeparableCNN.py:10-51."""

    def __init__(self, filters, num_of_patches, dropout=0.25, bias=False):
        super().__init__()
        if bias:
            raise NotImplementedError("fd_sepblock_fwd implements the reference's bias=False blocks")
        self.num_of_patches = num_of_patches
        self.pointwise_conv1 = nn.Conv2d(filters, filters, kernel_size=(1, 1), padding=0, bias=bias)
        self.depthwise_conv = nn.Conv2d(filters, filters, kernel_size=(3, 3), padding=1, groups=filters, bias=bias)
        self.pointwise_conv2 = nn.Conv2d(filters, filters, kernel_size=(1, 1), padding=0, bias=bias)
        self.max_pool = nn.MaxPool2d(2)
        self.leaky_relu = nn.LeakyReLU(0.2)
        self.dropout2d = nn.Dropout2d(dropout)

    def forward(self, x):  # pragma: no cover - the engine runs the block
        raise RuntimeError("ResidualBlock is executed by SeparableEngine (CUDA only); call the parent model")


class SeparableEngine:
    """Device state + kernel sequence of the separable backbone's forward: flat fp32 parameters (every nn.Parameter
    is a view), packed bf16 / tap-major weights, NHWC bf16 activation buffers per batch size."""

    def __init__(self, filters, in_ch, in_h, in_w, num_blocks, stem_k, stem_s, stem_pad, head_k, head_pad,
                 block_patches, slope=0.2):
        if filters != 64:
            raise NotImplementedError("the separable-block kernel is instantiated for 64 channels; got filters=%d"
                                      % filters)
        self.F, self.num_blocks, self.slope = filters, num_blocks, slope
        self.in_ch, self.in_h, self.in_w = in_ch, in_h, in_w
        self.stem_s, self.stem_pad, self.head_k, self.head_pad = stem_s, stem_pad, head_k, head_pad
        H = (in_h + 2 * stem_pad - stem_k) // stem_s + 1
        W = (in_w + 2 * stem_pad - stem_k) // stem_s + 1
        self.shapes, self.pools = [], []
        for _ in range(num_blocks):
            self.shapes.append((H, W))
            pool = H > block_patches                       # SeparableCNN.py:49
            self.pools.append(pool)
            if pool:
                H, W = H // 2, W // 2
        self.Hl, self.Wl = H, W
        self.So_h, self.So_w = H + 2 * head_pad - head_k + 1, W + 2 * head_pad - head_k + 1
        F_ = filters
        self.sections = [("conv1.weight", (F_, in_ch, stem_k, stem_k)), ("conv1.bias", (F_,)),
                         ("pw", (2 * num_blocks, F_, F_, 1, 1)), ("dw", (num_blocks, F_, 1, 3, 3)),
                         ("out.weight", (5, F_, head_k, head_k)), ("out.bias", (5,))]
        self.offsets, off = {}, 0
        for name, shape in self.sections:
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = (off, n, shape)
            off += (n + 3) // 4 * 4
        self.n_flat = off
        self.device = None
        self.pflat = None
        self.plans: Dict[int, dict] = {}

    def param_names(self):
        names = ["conv1.weight", "conv1.bias"]
        for k in range(self.num_blocks):
            names += [f"residual_blocks.{k}.pointwise_conv1.weight", f"residual_blocks.{k}.depthwise_conv.weight",
                      f"residual_blocks.{k}.pointwise_conv2.weight"]
        return names + ["out.weight", "out.bias"]

    def section(self, name):
        off, n, shape = self.offsets[name]
        return self.pflat[off:off + n].view(shape)

    def _view(self, name):
        if name.startswith("residual_blocks."):
            _, k, conv, _ = name.split(".")
            k = int(k)
            if conv == "depthwise_conv":
                return self.section("dw")[k]
            return self.section("pw")[2 * k + (0 if conv == "pointwise_conv1" else 1)]
        return self.section(name)

    def bind(self, params: Dict[str, nn.Parameter]):
        dev = params["conv1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("the fd_b200 backbone runs on CUDA only (no CPU fallback): call model.cuda()")
        if self.device != dev or self.pflat is None:
            self.device = dev
            self.pflat = torch.zeros(self.n_flat, dtype=F32, device=dev)
            nb, F_ = self.num_blocks, self.F
            self.w_pw = torch.empty((2 * nb, F_, F_), dtype=BF16, device=dev)
            self.w_dw = torch.empty((nb, 9, F_), dtype=F32, device=dev)
            self.w_head_t = torch.empty(self.head_k * self.head_k * 5 * F_, dtype=F32, device=dev)
            self.plans.clear()
        for name in self.param_names():
            p, v = params[name], self._view(name)
            if p.data_ptr() != v.data_ptr():
                with torch.no_grad():
                    v.copy_(p.data.to(device=dev, dtype=F32))
                p.data = v

    def plan(self, B):
        if B not in self.plans:
            def bf(h, w):
                return torch.empty((B, h, w, self.F), dtype=BF16, device=self.device)
            H0, W0 = self.shapes[0]
            pl = {"act0": bf(H0, W0), "out": []}
            for (h, w), pool in zip(self.shapes, self.pools):
                pl["out"].append(bf(h // 2, w // 2) if pool else bf(h, w))
            pl["y"] = torch.empty((B, 5, self.So_h, self.So_w), dtype=F32, device=self.device)
            self.plans[B] = pl
        return self.plans[B]

    def pack_weights(self):
        ops.sep_pack(self.section("pw"), self.w_pw, self.section("dw"), self.w_dw)
        ops.head_pack(self.section("out.weight"), self.w_head_t)

    def forward(self, x: torch.Tensor, repack: bool = True) -> torch.Tensor:
        """x: [B,in_ch,H,W] fp32 in [0,1] (or uint8: /255 fused into the stem).  Returns the sigmoid head
        [B,5,So,So] fp32 (a buffer of the plan, overwritten by the next call with the same batch size)."""
        pl = self.plan(x.shape[0])
        if repack:
            self.pack_weights()
        ops.stem_fwd(x, self.section("conv1.weight"), self.section("conv1.bias"), pl["act0"], self.stem_s, self.stem_pad)
        cur = pl["act0"]
        for k in range(self.num_blocks):
            # the MaxPool2d(2) of a pooling block is fused: the un-pooled sum never reaches HBM
            ops.sepblock_fwd(cur, self.w_pw[2 * k], self.w_dw[k], self.w_pw[2 * k + 1], self.slope, pl["out"][k],
                             pool=self.pools[k])
            cur = pl["out"][k]
        ops.head_fwd(cur, None, self.section("out.weight"), self.section("out.bias"), pl["y"], self.head_pad,
                     w_t=self.w_head_t)
        return pl["y"]


class SeparablePlanarEngine:
    """``filters = 64 * G`` (the reference's ``__main__`` builds 128, SeparableCNN.py:124) on G channel planes with the
    64-channel kernels: a pointwise 64G -> 64G convolution is, per output plane, a chain of G ``fd_conv3x3`` calls in
    centre-tap (``FD_CONV_1X1``) mode whose raw partial sums ride on the residual input (the skip connection is the
    first addend of the pw2 chain); LeakyReLU after pw1 in ``fd_act_mask``, depthwise 3x3 + LeakyReLU per plane in
    ``fd_dwconv3x3_lrelu``, pooling, stem and head (partial logits per plane) as in engine_planar.PlanarEngine.
    For G = 2 or 4 the pointwise convolutions run on ``fd_conv3x3_wide`` in centre-tap mode instead (one launch per
    convolution and group of 128 output channels, tcgen05.mma.cta_group::2, LeakyReLU / skip add fused: ``use_wide``).
    Functional path (the fused ``fd_sepblock_fwd`` is the 64-channel fast path); inference only."""

    def __init__(self, filters, in_ch, in_h, in_w, num_blocks, stem_k, stem_s, stem_pad, head_k, head_pad,
                 block_patches, slope=0.2):
        if filters % 64 != 0 or filters < 128:
            raise NotImplementedError("SeparableCNN kernels exist for filters = 64 (fused) and 64 * G (channel planes); "
                                      "got filters=%d" % filters)
        self.F, self.G, self.num_blocks, self.slope = filters, filters // 64, num_blocks, slope
        self.in_ch, self.in_h, self.in_w, self.stem_k = in_ch, in_h, in_w, stem_k
        self.stem_s, self.stem_pad, self.head_k, self.head_pad = stem_s, stem_pad, head_k, head_pad
        H = (in_h + 2 * stem_pad - stem_k) // stem_s + 1
        W = (in_w + 2 * stem_pad - stem_k) // stem_s + 1
        self.shapes, self.pools = [], []
        for _ in range(num_blocks):
            self.shapes.append((H, W))
            pool = H > block_patches
            self.pools.append(pool)
            if pool:
                H, W = H // 2, W // 2
        self.So_h, self.So_w = H + 2 * head_pad - head_k + 1, W + 2 * head_pad - head_k + 1
        self.device, self.params, self.plans = None, None, {}
        self.use_wide = self.G in (2, 4)
        self.w_wide = None

    def bind(self, params):
        dev = params["conv1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("the fd_b200 backbone runs on CUDA only (no CPU fallback): call model.cuda()")
        if dev != self.device:
            self.device = dev
            self.sides = [torch.cuda.Stream(device=dev) for _ in range(self.G - 1)]
            self.plans.clear()
        self.params = params

    # the per-plane chains of a block are independent: plane g > 0 runs on a side stream (parallel graph branches)
    def _fork(self):
        main = torch.cuda.current_stream()
        for s_ in self.sides:
            s_.wait_stream(main)
        return main

    def _join(self, main):
        for s_ in self.sides:
            main.wait_stream(s_)

    def _on(self, main, g):
        return torch.cuda.stream(main if g == 0 else self.sides[g - 1])

    def plan(self, B):
        if B not in self.plans:
            G, dev = self.G, self.device

            def planes(h, w):
                return [torch.empty((B, h, w, 64), dtype=BF16, device=dev) for _ in range(G)]
            H0, W0 = self.shapes[0]
            pl = {"act0": planes(H0, W0), "blocks": []}
            pl["x_cache"] = ops.stem_cache(B, (self.in_ch, self.in_h, self.in_w), (self.stem_k, self.stem_s, self.stem_pad),
                                           dev) if G > 1 else None
            for (h, w), pool in zip(self.shapes, self.pools):
                b = {"T": planes(h, w), "T2": planes(h, w), "t1": planes(h, w), "t2": planes(h, w)}
                b["s"] = planes(h, w)
                b["out"] = planes(h // 2, w // 2) if pool else b["s"]
                pl["blocks"].append(b)
            pl["y"] = torch.empty((B, 5, self.So_h, self.So_w), dtype=F32, device=dev)
            self.plans[B] = pl
        return self.plans[B]

    def _sub(self, layer, g, h):
        return (layer * self.G + g) * self.G + h

    def pack_weights(self):
        G, nb, P = self.G, self.num_blocks, self.params
        pw, dw = [], []
        for k in range(nb):
            pre = f"residual_blocks.{k}."
            pw += [P[pre + "pointwise_conv1.weight"].detach(), P[pre + "pointwise_conv2.weight"].detach()]
            dw.append(P[pre + "depthwise_conv.weight"].detach())
        L = 2 * nb
        if self.use_wide:       # [layer][group of 128 couts][input plane][tap][128][64]; only the centre tap is written / read
            if self.w_wide is None or self.w_wide.device != self.device:
                self.w_wide = torch.zeros((L, G // 2, G, 9, 128, 64), dtype=BF16, device=self.device)
            ops.pack_conv3x3_wide(torch.stack(pw).float().contiguous(), self.w_wide, None)
        else:
            sub = torch.stack(pw).float().view(L, G, 64, G, 64).permute(0, 1, 3, 2, 4).reshape(L * G * G, 64, 64)
            self.w_pk = torch.zeros((L * G * G, 9, 64, 64), dtype=BF16, device=self.device)      # only the centre tap is read
            self.w_pk[:, 4] = sub.to(BF16)
        self.w_dw = torch.stack(dw).float().view(nb, G, 64, 9).permute(0, 1, 3, 2).contiguous()     # [nb, G, 9, 64]

    def _pw_chain(self, srcs, layer, g, first_residual, dst_a, dst_b):
        prev = first_residual
        for h in range(self.G):
            dst = dst_b if (h % 2) else dst_a
            ops.conv3x3(srcs[h], self.w_pk[self._sub(layer, g, h)], slope=self.slope, lrelu=False, residual=prev, out=dst,
                        flags=ops.CONV_1X1)
            prev = dst
        return prev

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        G, nb, P = self.G, self.num_blocks, self.params
        pl = self.plan(x.shape[0])
        self.pack_weights()
        w1, b1 = P["conv1.weight"].detach().float(), P["conv1.bias"].detach().float()
        ops.stem_planes_fwd(x, w1, b1, pl["act0"], self.stem_s, self.stem_pad, x_cache=pl["x_cache"])
        cur = pl["act0"]
        for k, b in enumerate(pl["blocks"]):
            if self.use_wide:
                for go in range(G // 2):       # pw1 + LeakyReLU (models/SeparableCNN.py:42-43)
                    ops.conv3x3_wide(cur, self.w_wide[2 * k, go], slope=self.slope, lrelu=True,
                                     out=[b["t1"][2 * go], b["t1"][2 * go + 1]], flags=ops.CONV_1X1)
                main = self._fork()
                for g in range(G):
                    with self._on(main, g):
                        ops.dwconv3x3_lrelu(b["t1"][g], self.w_dw[k, g], self.slope, b["t2"][g])
                self._join(main)
                for go in range(G // 2):       # pw2 + skip (:46-48; Dropout2d is the identity in eval mode)
                    ops.conv3x3_wide(b["t2"], self.w_wide[2 * k + 1, go], slope=self.slope,
                                     residual=[cur[2 * go], cur[2 * go + 1]], out=[b["s"][2 * go], b["s"][2 * go + 1]],
                                     flags=ops.CONV_1X1)
                if self.pools[k]:
                    main = self._fork()
                    for g in range(G):
                        with self._on(main, g):
                            ops.maxpool2x2_fwd(b["s"][g], b["out"][g])
                    self._join(main)
                cur = b["out"]
                continue
            main = self._fork()
            for g in range(G):
                with self._on(main, g):
                    raw = self._pw_chain(cur, 2 * k, g, None, b["T"][g], b["T2"][g])
                    ops.act_mask(raw, self.slope, None, None, None, b["t1"][g])
                    ops.dwconv3x3_lrelu(b["t1"][g], self.w_dw[k, g], self.slope, b["t2"][g])
            self._join(main)
            main = self._fork()
            for g in range(G):
                with self._on(main, g):
                    # pw2 + skip: the block input is the first addend of the chain; the last call writes the sum
                    prev = cur[g]
                    for h in range(G):
                        dst = b["s"][g] if h == G - 1 else (b["T"][g] if (h % 2 == 0) else b["T2"][g])
                        ops.conv3x3(b["t2"][h], self.w_pk[self._sub(2 * k + 1, g, h)], slope=self.slope, lrelu=False,
                                    residual=prev, out=dst, flags=ops.CONV_1X1)
                        prev = dst
                    if self.pools[k]:
                        ops.maxpool2x2_fwd(b["s"][g], b["out"][g])
            self._join(main)
            cur = b["out"]
        wo = P["out.weight"].detach().float()
        logits = None
        for g in range(G):
            wg = wo[:, g * 64:(g + 1) * 64].contiguous()
            wt = torch.empty(self.head_k * self.head_k * 5 * 64, dtype=F32, device=x.device)
            ops.head_pack(wg, wt)
            part = torch.empty_like(pl["y"])
            ops.head_fwd(cur[g], None, wg, None, part, self.head_pad, w_t=wt)
            logits = part if logits is None else logits + part
        torch.sigmoid(logits + P["out.bias"].detach().float().view(1, 5, 1, 1), out=pl["y"])
        return pl["y"]


class _SeparableFn(torch.autograd.Function):
    """Autograd bridge: forward / backward of the whole separable backbone as one node (engine_separable)."""

    @staticmethod
    def forward(ctx, model, x, *params):
        eng = model.train_engine
        with torch.cuda.device(x.device):
            pl = eng.forward(x, dropout=model.training)
        ctx.eng, ctx.pl, ctx.generation = eng, pl, pl["generation"]
        return pl["y"].clone()

    @staticmethod
    def backward(ctx, dy):
        eng, pl = ctx.eng, ctx.pl
        if pl["generation"] != ctx.generation:
            raise RuntimeError("fd_b200 SeparableCNN: backward() after ANOTHER forward of the same batch size overwrote the "
                               "saved activations; call backward() before the next forward")
        with torch.cuda.device(dy.device):
            eng.run_backward(pl, dy.contiguous().float())
        return (None, None, *[eng.grad_view(n).clone() for n in eng.param_names()])


class SeparableCNN(BaseModel):
    def __init__(self, filters, input_shape, num_of_residual_blocks=10, probability_threshold=0.5, iou_threshold=0.5,
                 pretrained=False, input_kernel_size=10, input_stride=8, output_kernel_size=6, output_padding=0):
        super().__init__(filters, input_shape, num_of_patches=16,                    # SeparableCNN.py:69-75
                         probability_threshold=probability_threshold, iou_threshold=iou_threshold)
        self.pretrained = pretrained
        self.dropout2d = nn.Dropout2d(0.5)
        self.conv1 = nn.Conv2d(input_shape[0], filters, kernel_size=(input_kernel_size, input_kernel_size),
                               stride=(input_stride, input_stride), padding=input_kernel_size - input_stride)
        self.residual_blocks = nn.Sequential(
            *[ResidualBlock(filters=filters, num_of_patches=self.num_of_patches) for _ in range(num_of_residual_blocks)])
        self.out = nn.Conv2d(filters, 5, stride=(1, 1), kernel_size=(output_kernel_size, output_kernel_size),
                             padding=output_padding)
        self.sigmoid = nn.Sigmoid()
        Engine = SeparableEngine if filters == 64 else SeparablePlanarEngine       # fused kernel / channel planes
        eargs = (filters, input_shape[0], input_shape[1], input_shape[2], num_of_residual_blocks, input_kernel_size,
                 input_stride, input_kernel_size - input_stride, output_kernel_size, output_padding)
        self.engine = Engine(*eargs, block_patches=self.num_of_patches)                       # inference
        from ..engine_separable import SeparableTrainEngine
        self.train_engine = SeparableTrainEngine(*eargs, block_patches=self.num_of_patches)   # forward with saved state + backward

    def forward(self, x: torch.Tensor, predict: torch.Tensor = torch.tensor(0)):
        is_predict = bool(predict == 1)
        if is_predict:                                     # SeparableCNN.py:105-108
            x = self._resize(self._to_model_device(x))
            if x.dtype != torch.uint8:
                x = x / 255.0
            if len(x.shape) == 3:
                x = torch.unsqueeze(x, 0)
        if not x.is_cuda:
            raise RuntimeError("fd_b200 models run on CUDA tensors only (no CPU fallback)")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        x = x.contiguous()
        params = dict(self.named_parameters())
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params.values())
        if needs_grad or self.training:
            # training path (Dropout2d active in train mode, activations saved): layer-by-layer engine
            self.train_engine.bind(params)
            if needs_grad:
                y = _SeparableFn.apply(self, x, *[params[n] for n in self.train_engine.param_names()])
            else:
                with torch.cuda.device(x.device):
                    y = self.train_engine.forward(x, dropout=True)["y"].clone()
        else:
            self.engine.bind(params)
            with torch.no_grad():
                y = self.engine.forward(x).clone()
        if is_predict:
            return self.single_non_max_suppression(y[0])   # SeparableCNN.py:114-115
        return y

    def train_step(self, x: torch.Tensor, gt: torch.Tensor, optimizer=None, allreduce=None):
        """forward + summed YoloLoss + backward in one call sequence (models/ModelMeta.py:141,173-176); returns the summed
        loss, gradients land in ``p.grad`` (views of ``self.train_engine.gflat``)."""
        params = dict(self.named_parameters())
        self.train_engine.bind(params)
        if not x.is_cuda:
            raise RuntimeError("fd_b200 models run on CUDA tensors only (no CPU fallback)")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        pl = self.train_engine.train_step(x.contiguous(), gt.float().contiguous(), dropout=self.training,
                                          allreduce=allreduce, optimizer=optimizer)
        for n, p in params.items():
            p.grad = self.train_engine.grad_view(n)
        return pl["loss"].sum()
