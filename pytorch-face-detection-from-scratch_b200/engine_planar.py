"""Residual grid backbones wider than the 64-channel tensor-core kernels (``filters = 128``: the width
``train_model.py:16`` trains) on G = filters / 64 channel PLANES.

Every activation / gradient tensor is stored as G separate NHWC tensors of 64 channels, so that the sm_100a kernels
instantiated for 64 channels (``fd_conv3x3``, ``fd_conv3x3_wgrad``, ``fd_stem_*``, ``fd_maxpool2x2_*``) are reused as
they are:

* a 64G -> 64G 3x3 convolution is, per output plane g, the SUM over input planes h of 64 -> 64 convolutions with the
  weight sub-block ``W[64g:64g+64, 64h:64h+64]``.  The sum is chained through the conv kernel's raw residual add (second
  call: ``residual`` = the first call's output, no activation), and the activation of models/PoolResnet.py:35-40 runs
  in ``fd_act_mask`` (LeakyReLU + sign-bit mask + Dropout2d multiplier + skip) / ``fd_grad_mask`` (its backward);
* the weight gradient of sub-block (g, h) is ``fd_conv3x3_wgrad(x plane h, gradient plane g)``;
* the stem runs once per output plane (the bf16 image cache is written by the first launch), the head on the
  64-channel head kernels plane by plane (partial logits summed before the sigmoid; tensor-core backward per plane).

For G = 2 or 4 the forward and input-gradient convolutions run on the NATIVE wide kernel instead (``fd_conv3x3_wide``:
one launch per layer and group of 128 output channels, tcgen05.mma.cta_group::2 with N = 128, partial sums over the input
planes in TMEM, activation / mask / dropout / skip fused into its epilogue -- ``use_wide``); the chained 64-channel path
above stays as the fallback for other widths (one more bf16 rounding per convolution for the chained partial sum).  The
weight gradients still run per (gradient plane, input plane) pair on ``fd_conv3x3_wgrad_multi``.
Parameters stay ordinary ``nn.Parameter`` tensors; gradients land in one flat fp32 buffer (``gflat``, the all-reduce
unit) of which every ``p.grad`` is a view.
"""
from __future__ import annotations

from typing import Callable, Dict, List

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


class _PBlock:
    __slots__ = ("H", "W", "pool", "inp", "T", "T2", "a", "ma", "mb", "s", "out", "amax", "G", "gs", "gp1", "gp2", "U")


class _PPlan:
    def __init__(self, eng: "PlanarEngine", B: int, train: bool, device):
        G = eng.G
        self.B, self.train = B, train

        def planes(h, w):
            return [torch.empty((B, h, w, 64), dtype=BF16, device=device) for _ in range(G)]

        def masks(h, w):
            return [torch.empty((B, h, w, 2), dtype=torch.int32, device=device) for _ in range(G)]

        # Runs of equal-shape blocks (a pooled block is a run of its own).  Per run and plane the weight-gradient
        # operands are STACKED and interleaved like in BackboneEngine -- XA[2i] = input of block k0+i, XA[2i+1] = its a,
        # GP[2i] = gp1, GP[2i+1] = gp2 -- so that ONE multi-problem launch per (gradient plane, input plane) covers all
        # 2n layers of the run (12 weight-gradient launches per step instead of 80).
        shapes, H, W = [], eng.H0, eng.W0
        for k in range(eng.num_blocks):
            shapes.append((H, W))
            if eng.pools[k]:
                H, W = H // 2, W // 2
        runs, k = [], 0
        while k < eng.num_blocks:
            k1 = k
            while (not eng.pools[k1]) and k1 + 1 < eng.num_blocks and shapes[k1 + 1] == shapes[k]:
                k1 += 1
            runs.append((k, k1))
            k = k1 + 1
        self.runs = runs
        run_of, self.XA, self.GP = {}, {}, {}
        for (k0, k1) in runs:
            n = k1 - k0 + 1
            h, w = shapes[k0]
            self.XA[k0] = [torch.empty((2 * n, B, h, w, 64), dtype=BF16, device=device) for _ in range(G)]
            self.GP[k0] = [torch.empty((2 * n, B, h, w, 64), dtype=BF16, device=device) for _ in range(G)] if train else None
            for kk in range(k0, k1 + 1):
                run_of[kk] = k0

        def input_planes(kk):
            """Planes holding the INPUT of block kk (= output of block kk-1 / the stem)."""
            k0 = run_of[kk]
            return [self.XA[k0][g][2 * (kk - k0)] for g in range(G)]

        H, W = eng.H0, eng.W0
        self.act0 = input_planes(0)
        self.blocks: List[_PBlock] = []
        cur = self.act0
        for k in range(eng.num_blocks):
            b = _PBlock()
            b.H, b.W, b.pool = H, W, eng.pools[k]
            b.inp = cur
            k0 = run_of[k]
            i = k - k0
            b.T, b.T2 = (planes(H, W), planes(H, W)) if not eng.use_wide else (None, None)
            b.a = [self.XA[k0][g][2 * i + 1] for g in range(G)]
            b.ma = masks(H, W) if train else [None] * G
            b.mb = masks(H, W) if train else [None] * G
            nxt = input_planes(k + 1) if k + 1 < eng.num_blocks else None
            b.amax = [None] * G
            if b.pool:
                b.s = planes(H, W)
                H, W = H // 2, W // 2
                b.out = nxt if nxt is not None else planes(H, W)
                if train:
                    b.amax = [torch.empty((B, H, W, 8), dtype=torch.int16, device=device) for _ in range(G)]
            else:
                b.s = nxt if nxt is not None else planes(H, W)
                b.out = b.s
            if train:
                b.G = planes(H, W)                      # gradient w.r.t. the block OUTPUT (post-pool resolution)
                b.gs = planes(b.H, b.W) if b.pool else None
                b.gp1 = [self.GP[k0][g][2 * i] for g in range(G)]
                b.gp2 = [self.GP[k0][g][2 * i + 1] for g in range(G)]
                b.U = planes(b.H, b.W) if not eng.use_wide else None
            self.blocks.append(b)
            cur = b.out
        self.y = torch.empty((B, 5, eng.So_h, eng.So_w), dtype=F32, device=device)
        # bf16 image copy: read by the stem weight gradient and by the forward of every plane after the first
        self.x_cache = None
        if train or G > 1:
            self.x_cache = ops.stem_cache(B, (eng.in_ch, eng.in_h, eng.in_w), (eng.stem_k, eng.stem_s, eng.stem_pad), device)
        if train:
            self.g_stem = planes(eng.H0, eng.W0)
            self.loss = torch.empty((B,), dtype=F32, device=device)
            self.dy = torch.empty_like(self.y)
        self.drop = None
        self.generation = 0
        self.x = self.head_in = self.w_head = None


class PlanarEngine:
    def __init__(self, filters: int, in_ch: int, in_h: int, in_w: int, num_blocks: int, stem_k: int, stem_s: int,
                 stem_pad: int, head_k: int, head_pad: int, pool_rule: Callable[[int], bool], slope: float = 0.2,
                 block_drop: float = 0.25, head_drop: float = 0.5):
        if filters % 64 != 0 or filters < 128:
            raise NotImplementedError("PlanarEngine handles filters = 64 * G with G >= 2; got %d" % filters)
        self.F, self.G = filters, filters // 64
        self.in_ch, self.in_h, self.in_w = in_ch, in_h, in_w
        self.num_blocks, self.slope = num_blocks, slope
        self.stem_k, self.stem_s, self.stem_pad, self.head_k, self.head_pad = stem_k, stem_s, stem_pad, head_k, head_pad
        self.block_drop, self.head_drop = block_drop, head_drop
        self.H0 = (in_h + 2 * stem_pad - stem_k) // stem_s + 1
        self.W0 = (in_w + 2 * stem_pad - stem_k) // stem_s + 1
        self.pools = []
        H, W = self.H0, self.W0
        for _ in range(num_blocks):
            p = bool(pool_rule(H))
            self.pools.append(p)
            if p:
                H, W = H // 2, W // 2
        self.Hl, self.Wl = H, W
        self.So_h, self.So_w = H + 2 * head_pad - head_k + 1, W + 2 * head_pad - head_k + 1
        F_ = filters
        self.sections = [("conv1.weight", (F_, in_ch, stem_k, stem_k)), ("conv1.bias", (F_,)),
                         ("w3", (2 * num_blocks, F_, F_, 3, 3)), ("b3", (2 * num_blocks, F_)),
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
        self.gflat = None
        self.use_wide = self.G in (2, 4)      # fd_conv3x3_wide: groups of 128 output channels, <= 4 input planes
        self.plans: Dict[tuple, _PPlan] = {}

    # ------------------------------------------------------------------ parameters / gradients
    def param_names(self) -> List[str]:
        names = ["conv1.weight", "conv1.bias"]
        for k in range(self.num_blocks):
            for c in ("conv1", "conv2"):
                names += [f"residual_blocks.{k}.{c}.weight", f"residual_blocks.{k}.{c}.bias"]
        return names + ["out.weight", "out.bias"]

    def section(self, flat, name):
        off, n, shape = self.offsets[name]
        return flat[off:off + n].view(shape)

    def _view(self, flat, name: str) -> torch.Tensor:
        if name.startswith("residual_blocks."):
            _, k, c, kind = name.split(".")
            layer = 2 * int(k) + (0 if c == "conv1" else 1)
            return self.section(flat, "w3" if kind == "weight" else "b3")[layer]
        return self.section(flat, name)

    def grad_view(self, name: str) -> torch.Tensor:
        return self._view(self.gflat, name)

    # flat fp32 parameter / gradient buffers: the unit of the one-kernel Adam (optim.FlatAdam) and of the data-parallel exchange
    def opt_params(self) -> torch.Tensor:
        return self.pflat

    def opt_grads(self) -> torch.Tensor:
        return self.gflat

    def bind(self, params):
        dev = params["conv1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("the fd_b200 backbone runs on CUDA only (no CPU fallback): call model.cuda()")
        if self.device != dev or self.gflat is None:
            self.device = dev
            G, L = self.G, 2 * self.num_blocks
            self.gflat = torch.zeros(self.n_flat, dtype=F32, device=dev)
            self.pflat = torch.zeros(self.n_flat, dtype=F32, device=dev)
            self.dwp = torch.zeros((L * G * G, 9 * 64 * 64), dtype=F32, device=dev)        # packed accumulators
            self.gb3 = self.section(self.gflat, "b3").view(L, G, 64)       # bias gradients accumulate in place (gflat is zeroed per step)
            self.w_fwd = torch.empty((L * G * G, 9, 64, 64), dtype=BF16, device=dev)
            self.w_dgrad = torch.empty((L * G * G, 9, 64, 64), dtype=BF16, device=dev)
            if self.use_wide:             # [layer][group of 128 couts][input plane][tap][128][64]
                self.w_fwd_wide = torch.empty((L, G // 2, G, 9, 128, 64), dtype=BF16, device=dev)
                self.w_dgrad_wide = torch.empty((L, G // 2, G, 9, 128, 64), dtype=BF16, device=dev)
            self.sides = [torch.cuda.Stream(device=dev) for _ in range(self.G - 1)]
            self.plans.clear()
        # every nn.Parameter becomes a view of the flat buffer (values preserved): the stacked [L,F,F,3,3] weights the pack
        # kernel reads and the buffer FlatAdam updates are then the parameters themselves
        for name in self.param_names():
            prm = params[name]
            v = self._view(self.pflat, name)
            if prm.data_ptr() != v.data_ptr():
                with torch.no_grad():
                    v.copy_(prm.data.to(device=dev, dtype=F32))
                prm.data = v
        self.params = params

    def plan(self, B, train):
        key = (B, train)
        if key not in self.plans:
            self.plans[key] = _PPlan(self, B, train, self.device)
        return self.plans[key]

    def _sub(self, layer, g, h):
        return (layer * self.G + g) * self.G + h

    # The per-plane chains of a layer are independent of each other and, behind the last pooling stage, each of their
    # launches fills less than half of the GPU (one tile per image): plane g > 0 runs on a side stream, forked from and
    # joined into the main stream around every group of chains (captured into the CUDA graph as parallel branches).
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

    def pack_weights(self):
        """[L,64G,64G,3,3] fp32 -> sub-blocks [(L,g,h),64,64,3,3] (torch data movement) -> bf16 forward / dgrad packing."""
        G, L = self.G, 2 * self.num_blocks
        w_all = self.section(self.pflat, "w3")          # [L,F,F,3,3]: the parameters themselves (bind)
        if self.use_wide:
            ops.pack_conv3x3_wide(w_all, self.w_fwd_wide, self.w_dgrad_wide)
        else:
            w3 = w_all.view(L, G, 64, G, 64, 3, 3).permute(0, 1, 3, 2, 4, 5, 6).contiguous()
            ops.pack_conv3x3(w3.view(L * G * G, 64, 64, 3, 3), self.w_fwd, self.w_dgrad)
        self.b3 = self.section(self.pflat, "b3").view(L, G, 64)

    # ------------------------------------------------------------------ forward
    def _conv_sum(self, srcs, wsel, layer, g, bias, dst_a, dst_b, **last_kw):
        """sum_h conv(srcs[h], W[layer][g][h]) (+ bias) as a chain of 64-channel convolutions; returns the raw sum's
        buffer (or None when the last call wrote a masked out2 via last_kw)."""
        prev = None
        for h in range(self.G):
            last = h == self.G - 1
            dst = dst_b if (h % 2) else dst_a
            kw = dict(bias=bias if h == 0 else None, slope=self.slope, lrelu=False, residual=prev)
            if last and last_kw:
                kw.update(last_kw)
            else:
                kw["out"] = dst
            ops.conv3x3(srcs[h], wsel[self._sub(layer, g, h)], **kw)
            prev = dst
        return prev

    def _pairs(self, planes, go):
        return None if planes is None or planes[0] is None else [planes[2 * go], planes[2 * go + 1]]

    kChainBlocks = 8        # fd_conv3x3_wide_chain takes up to 16 layers

    def _chain_runs(self, pl):
        """{k0: k1} of the runs of >= 2 blocks whose maps fit a CTA pair (fd_conv3x3_wide_chain): 128-filter models only."""
        if not (self.use_wide and self.G == 2):
            return {}
        return {k0: k1 for (k0, k1) in pl.runs
                if k1 > k0 and ops.conv3x3_wide_chain_ok(pl.B, pl.blocks[k0].H, pl.blocks[k0].W)}

    def _run_forward_chain(self, pl, k0, k1):
        """models/PoolResnet.py:35-42 for the blocks k0..k1 of one run: 2 (k1 - k0 + 1) convolutions, one launch per 8 blocks."""
        for c0 in range(k0, k1 + 1, self.kChainBlocks):
            layers = []
            for k in range(c0, min(c0 + self.kChainBlocks, k1 + 1)):
                blk, i = pl.blocks[k], k - k0
                drop = [pl.drop[k, g] for g in range(2)] if pl.drop is not None else None
                layers.append(ops.wide_chain_layer(2 * i, 2 * k, bias=self.b3[2 * k].reshape(-1), lrelu=True,
                                                   mask_out=self._pairs(blk.ma, 0), out=blk.a))
                layers.append(ops.wide_chain_layer(2 * i + 1, 2 * k + 1, bias=self.b3[2 * k + 1].reshape(-1), lrelu=True,
                                                   chan_scale=drop, residual=blk.inp, mask_out=self._pairs(blk.mb, 0), out=blk.s))
            ops.conv3x3_wide_chain(pl.XA[k0], self.w_fwd_wide, layers, slope=self.slope)
        last = pl.blocks[k1]
        if last.pool:
            main = self._fork()
            for g in range(self.G):
                with self._on(main, g):
                    ops.maxpool2x2_fwd(last.s[g], last.out[g], last.amax[g])
            self._join(main)

    def _run_dgrad_chain(self, pl, k0, k1, drop, fused_gp2):
        """Input-gradient pass of the run k0..k1 (gp2 of block k1 is final): per block gp1 = dgrad(gp2, W2) * lrelu'(a), then
        G_{k-1} = dgrad(gp1, W1) + GS together with the previous block's gp2 = G_{k-1} * lrelu'(b) * dropout."""
        k = k1
        while k >= k0:
            c0 = max(k0, k - self.kChainBlocks + 1)
            layers = []
            for kk in range(k, c0 - 1, -1):
                blk, i = pl.blocks[kk], kk - k0
                GS = blk.gs if blk.pool else blk.G
                gprev = pl.blocks[kk - 1].G if kk > 0 else pl.g_stem
                layers.append(ops.wide_chain_layer(2 * i + 1, 2 * kk + 1, mask_in=blk.ma, out2=blk.gp1))
                if kk > k0:
                    prev = pl.blocks[kk - 1]
                    cs2 = [drop[kk - 1, 0], drop[kk - 1, 1]] if drop is not None else None
                    layers.append(ops.wide_chain_layer(2 * i, 2 * kk, residual=GS, out=gprev, mask_in=prev.mb, chan_scale2=cs2,
                                                       out2=prev.gp2))
                    fused_gp2.add(kk - 1)
                else:
                    layers.append(ops.wide_chain_layer(2 * i, 2 * kk, residual=GS, out=gprev))
            ops.conv3x3_wide_chain(pl.GP[k0], self.w_dgrad_wide, layers, slope=self.slope)
            k = c0 - 1

    def _block_forward_wide(self, pl, k, blk, cur):
        """models/PoolResnet.py:35-42 for one block on fd_conv3x3_wide: one launch per convolution and group of 128 couts."""
        G = self.G
        drop = [pl.drop[k, g] for g in range(G)] if pl.drop is not None else None
        for go in range(G // 2):
            ops.conv3x3_wide(cur, self.w_fwd_wide[2 * k, go], bias=self.b3[2 * k, 2 * go:2 * go + 2].reshape(-1),
                             slope=self.slope, lrelu=True, mask_out=self._pairs(blk.ma, go), out=self._pairs(blk.a, go))
        for go in range(G // 2):
            ops.conv3x3_wide(blk.a, self.w_fwd_wide[2 * k + 1, go], bias=self.b3[2 * k + 1, 2 * go:2 * go + 2].reshape(-1),
                             slope=self.slope, lrelu=True, chan_scale=self._pairs(drop, go), residual=self._pairs(cur, go),
                             mask_out=self._pairs(blk.mb, go), out=self._pairs(blk.s, go))
        if blk.pool:
            main = self._fork()
            for g in range(G):
                with self._on(main, g):
                    ops.maxpool2x2_fwd(blk.s[g], blk.out[g], blk.amax[g])
            self._join(main)

    def forward(self, x, train: bool, dropout: bool = False):
        B = x.shape[0]
        pl = self.plan(B, train)
        pl.generation += 1
        G, nb, P = self.G, self.num_blocks, self.params
        self.pack_weights()
        if dropout:
            r = torch.rand((nb + 1, G, B, 64), device=x.device)
            scale = torch.empty_like(r)
            ops.dropout_scale(r, nb * G * B * 64, 1.0 - self.block_drop, 1.0 - self.head_drop, scale)
            pl.drop = scale
        else:
            pl.drop = None
        pl.x = x
        w1, b1 = P["conv1.weight"].detach().float(), P["conv1.bias"].detach().float()
        ops.stem_planes_fwd(x, w1, b1, pl.act0, self.stem_s, self.stem_pad, x_cache=pl.x_cache)
        cur = pl.act0
        chain_end = self._chain_runs(pl)
        skip_until = -1
        for k, blk in enumerate(pl.blocks):
            if self.use_wide:
                if k in chain_end:                  # the whole run k .. chain_end[k] in one launch per 8 blocks
                    self._run_forward_chain(pl, k, chain_end[k])
                    skip_until = chain_end[k]
                if k > skip_until:
                    self._block_forward_wide(pl, k, blk, cur)
                cur = blk.out
                continue
            main = self._fork()
            for g in range(G):
                with self._on(main, g):
                    raw = self._conv_sum(cur, self.w_fwd, 2 * k, g, self.b3[2 * k, g], blk.T[g], blk.T2[g])
                    ops.act_mask(raw, self.slope, None, None, blk.ma[g], blk.a[g])
            self._join(main)
            main = self._fork()
            for g in range(G):
                with self._on(main, g):
                    raw = self._conv_sum(blk.a, self.w_fwd, 2 * k + 1, g, self.b3[2 * k + 1, g], blk.T[g], blk.T2[g])
                    ops.act_mask(raw, self.slope, pl.drop[k, g] if pl.drop is not None else None, cur[g], blk.mb[g],
                                 blk.s[g])
                    if blk.pool:
                        ops.maxpool2x2_fwd(blk.s[g], blk.out[g], blk.amax[g])
            self._join(main)
            cur = blk.out
        # head: partial logits per plane through the 64-channel kernel (bias=None), summed + bias + sigmoid on [B,5,S,S]
        wo = P["out.weight"].detach().float()
        pl.head_in = cur
        pl.w_head = [wo[:, g * 64:(g + 1) * 64].contiguous() for g in range(G)]
        # the per-plane head kernels fill less than half of the SMs each: run them side by side (buffers from the main stream)
        wts = [torch.empty(self.head_k * self.head_k * 5 * 64, dtype=F32, device=x.device) for _ in range(G)]
        parts = [torch.empty_like(pl.y) for _ in range(G)]
        main = self._fork()
        for g in range(G):
            with self._on(main, g):
                ops.head_pack(pl.w_head[g], wts[g])
                ops.head_fwd(cur[g], pl.drop[nb, g] if pl.drop is not None else None, pl.w_head[g], None, parts[g], self.head_pad,
                             w_t=wts[g])
        self._join(main)
        logits = parts[0]
        for g in range(1, G):
            logits = logits + parts[g]
        torch.sigmoid(logits + P["out.bias"].detach().float().view(1, 5, 1, 1), out=pl.y)
        return pl

    # ------------------------------------------------------------------ backward
    def run_backward(self, pl: _PPlan, dy: torch.Tensor):
        assert pl.train
        G, nb, P = self.G, self.num_blocks, self.params
        drop = pl.drop
        self.gflat.zero_()
        self.dwp.zero_()
        last = pl.blocks[nb - 1]
        gw_out = self.section(self.gflat, "out.weight")
        # tensor-core head backward per plane, side by side; dbias is the same for every plane: count it once
        dwgs = [torch.zeros_like(pl.w_head[g]) for g in range(G)]
        dbgs = [self.section(self.gflat, "out.bias") if g == 0 else torch.zeros(5, dtype=F32, device=dy.device) for g in range(G)]
        main = self._fork()
        for g in range(G):
            with self._on(main, g):
                ops.head_bwd(pl.head_in[g], drop[nb, g] if drop is not None else None, pl.w_head[g], pl.y, dy, self.head_pad,
                             last.G[g], None, None, self.slope, None, dwgs[g], dbgs[g])
                gw_out[:, g * 64:(g + 1) * 64].copy_(dwgs[g])
        self._join(main)
        fused_gp2 = set()
        chain_start = {k1: k0 for k0, k1 in self._chain_runs(pl).items()}
        chained = set()
        for k in range(nb - 1, -1, -1):
            blk = pl.blocks[k]
            L1, L2 = 2 * k, 2 * k + 1
            if not (k in fused_gp2):        # gp2 of this block was already written by the previous iteration's dual-output launch
                main = self._fork()
                for g in range(G):
                    with self._on(main, g):
                        cs = drop[k, g] if drop is not None else None
                        if blk.pool:
                            ops.maxpool2x2_bwd(blk.s[g], blk.G[g], blk.gs[g], blk.mb[g], cs, self.slope, blk.gp2[g],
                                               argmax=blk.amax[g])
                        else:
                            ops.grad_mask(blk.G[g], self.slope, blk.mb[g], cs, blk.gp2[g])
                self._join(main)
            GS = blk.gs if blk.pool else blk.G
            gprev = pl.blocks[k - 1].G if k > 0 else pl.g_stem
            if k in chain_start:                # last block of a chained run: its gp2 is final -> the run's whole dgrad pass
                self._run_dgrad_chain(pl, chain_start[k], k, drop, fused_gp2)
                chained.update(range(chain_start[k], k + 1))
            if k in chained:
                pass
            elif self.use_wide:
                # gp1 = dgrad(gp2, W2) * lrelu'(a);  G_{k-1} = dgrad(gp1, W1) + GS  -- one launch per group of 128 channels
                for go in range(G // 2):
                    ops.conv3x3_wide(blk.gp2, self.w_dgrad_wide[L2, go], slope=self.slope, mask_in=self._pairs(blk.ma, go),
                                     out2=self._pairs(blk.gp1, go))
                # the previous block's conv2 gradient (gp2 = G * lrelu'(b) * dropout) rides on the same launch when the map
                # runs in shared-tile mode and the previous block does not pool (its G is exactly this launch's output)
                prev = pl.blocks[k - 1] if k > 0 else None
                dual = prev is not None and not prev.pool and ops.conv3x3_wide_shared_tile(pl.B, blk.H, blk.W)
                for go in range(G // 2):
                    if dual:
                        cs2 = [drop[k - 1, 2 * go], drop[k - 1, 2 * go + 1]] if drop is not None else None
                        ops.conv3x3_wide(blk.gp1, self.w_dgrad_wide[L1, go], slope=self.slope, residual=self._pairs(GS, go),
                                         out=self._pairs(gprev, go), mask_in=self._pairs(prev.mb, go), chan_scale2=cs2,
                                         out2=self._pairs(prev.gp2, go))
                    else:
                        ops.conv3x3_wide(blk.gp1, self.w_dgrad_wide[L1, go], slope=self.slope, residual=self._pairs(GS, go),
                                         out=self._pairs(gprev, go))
                if dual:
                    fused_gp2.add(k - 1)
            else:
                self._block_dgrad_planes(blk, L1, L2, GS, gprev)
            # weight / bias gradients: once per run, when the gp1 / gp2 of all its blocks are final
            if k in pl.XA:
                n3 = 9 * 64 * 64
                dwp_flat, gb3_flat = self.dwp.view(-1), self.gb3.view(-1)
                if self.use_wide:       # one 128 x 128 channel block of dW per call (fd_conv3x3_wgrad_wide)
                    for gg in range(G // 2):
                        for hh in range(G // 2):
                            sub_off = [self._sub(2 * k, 2 * gg + c, 2 * hh + r) * n3 for r in range(2) for c in range(2)]
                            ops.conv3x3_wgrad_wide(pl.XA[k][2 * hh], pl.XA[k][2 * hh + 1], pl.GP[k][2 * gg], pl.GP[k][2 * gg + 1],
                                                   dwp_flat, sub_off, dw_stride=G * G * n3,
                                                   dbias0=gb3_flat[(2 * k * G + 2 * gg) * 64:] if hh == 0 else None,
                                                   dbias1=gb3_flat[(2 * k * G + 2 * gg + 1) * 64:] if hh == 0 else None,
                                                   dbias_stride=G * 64)
                    continue
                for g in range(G):
                    for h in range(G):
                        first = self._sub(2 * k, g, h)
                        ops.conv3x3_wgrad_multi(pl.XA[k][h], pl.GP[k][g], dwp_flat[first * n3:], G * G * n3,
                                                gb3_flat[(2 * k * G + g) * 64:] if h == 0 else None, G * 64)
        gw1, gb1 = self.section(self.gflat, "conv1.weight"), self.section(self.gflat, "conv1.bias")
        ops.stem_planes_wgrad(pl.x, pl.g_stem, gw1, gb1, self.stem_s, self.stem_pad, x_cache=pl.x_cache)
        L = 2 * nb
        # packed sub-blocks -> the [L,F,F,3,3] gradient section in one pass; the bias gradients were accumulated in place
        ops.unpack_wgrad3x3_planes(self.dwp.view(L * G * G, 9, 64, 64), G, self.section(self.gflat, "w3"))

    def _block_dgrad_planes(self, blk, L1, L2, GS, gprev):
        G = self.G
        # gp1[h] = (sum_g dgrad(gp2[g], W2[g][h])) * lrelu'(a[h])
        main = self._fork()
        for h in range(G):
            with self._on(main, h):
                prev = None
                for g in range(G):
                    w = self.w_dgrad[self._sub(L2, g, h)]
                    if g < G - 1:
                        dst = blk.U[h] if (g % 2 == 0) else blk.T[h]
                        ops.conv3x3(blk.gp2[g], w, slope=self.slope, residual=prev, out=dst)
                        prev = dst
                    else:
                        ops.conv3x3(blk.gp2[g], w, slope=self.slope, residual=prev, mask_in=blk.ma[h],
                                    out2=blk.gp1[h])
        self._join(main)
        # G_{k-1}[h] = sum_g dgrad(gp1[g], W1[g][h]) + GS[h]
        main = self._fork()
        for h in range(G):
            with self._on(main, h):
                prev = GS[h]
                for g in range(G):
                    w = self.w_dgrad[self._sub(L1, g, h)]
                    dst = gprev[h] if g == G - 1 else (blk.U[h] if (g % 2 == 0) else blk.T[h])
                    ops.conv3x3(blk.gp1[g], w, slope=self.slope, residual=prev, out=dst)
                    prev = dst
        self._join(main)

    def train_step(self, x, gt, dropout: bool = True, allreduce=None, optimizer=None):
        pl = self.forward(x, train=True, dropout=dropout)
        if tuple(gt.shape) != tuple(pl.y.shape) or gt.device != pl.y.device:
            raise ValueError(f"target map {tuple(gt.shape)} on {gt.device} does not match the head "
                             f"{tuple(pl.y.shape)} on {pl.y.device}")
        ops.yolo_loss(pl.y, gt, pl.loss, None, pl.dy)
        self.run_backward(pl, pl.dy)
        if allreduce is not None:
            allreduce(self.gflat)
        if optimizer is not None:
            optimizer.step()
        return pl
