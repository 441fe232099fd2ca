"""Execution engine of the residual grid backbones (PoolResnet / Resnet of the reference).

Owns the device-resident state the C-ABI kernels work on -- one flat fp32 parameter buffer, one
flat fp32 gradient buffer (the all-reduce unit for data parallelism), bf16 packed weights, and
per-batch-size activation plans in NHWC bf16 -- and issues the kernel sequence of a forward pass
and of the backward pass.  PyTorch is used for memory, streams and CUDA graphs only.

Reference call sites replaced: models/PoolResnet.py:93-105 (forward), the autograd backward of the
same graph, and models/ModelMeta.py:141,173-176 (train step = forward + summed YoloLoss + backward).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


class _Block:
    __slots__ = ("H", "W", "pool", "a", "ma", "mb", "s", "out", "G", "gs", "gp1", "gp2", "amax")


class _Plan:
    """Activation / gradient buffers for one (batch size, train?) combination."""

    def __init__(self, eng: "BackboneEngine", B: int, train: bool, device):
        F = eng.F
        self.B, self.train = B, train

        def bf(h, w):
            return torch.empty((B, h, w, F), dtype=BF16, device=device)

        # Blocks are grouped into maximal runs of equal (pre-pool) shape -- a pooled block is a run of its own.  Per
        # run the operands of the weight gradients are allocated STACKED and INTERLEAVED:
        #   XA[2i] = input of block k0+i, XA[2i+1] = its conv1 activation a;  GP[2i] = gp1, GP[2i+1] = gp2
        # so that problem j of ONE multi-problem wgrad launch is layer 2*k0 + j (uniform strides in dw / dbias).
        self.groups = []            # (k0, k1, XA_all, GP_all)
        self.chains = {}            # k0 -> chain record for runs executed by the fused chain kernels
        H, W = eng.H0, eng.W0
        shapes = []
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
        run_of = {}
        for (k0, k1) in runs:
            for k in range(k0, k1 + 1):
                run_of[k] = (k0, k1)

        def stack(n, h, w):
            return torch.empty((n, B, h, w, F), dtype=BF16, device=device)

        stacks = {}
        for (k0, k1) in runs:
            n = k1 - k0 + 1
            h, w = shapes[k0]
            stacks[k0] = {"XA": stack(2 * n, h, w), "GP": stack(2 * n, h, w) if train else None}
            self.groups.append((k0, k1, stacks[k0]["XA"], stacks[k0]["GP"]))
            if n > 1 and eng.use_chain and ops.resblock_chain_ok(h, w, F):
                # the whole run executes as ONE persistent kernel per direction (csrc/resblock_chain.cu)
                self.chains[k0] = {"k0": k0, "k1": k1}

        def out_buffer(k, h, w):
            """Buffer holding the OUTPUT of block k (k = -1: the stem) = input of block k+1."""
            nxt = k + 1
            if nxt in run_of:
                k0, _ = run_of[nxt]
                return stacks[k0]["XA"][2 * (nxt - k0)]
            return bf(h, w)

        H, W = eng.H0, eng.W0
        self.act0 = out_buffer(-1, H, W)
        self.blocks: List[_Block] = []
        for k in range(eng.num_blocks):
            blk = _Block()
            blk.H, blk.W, blk.pool = H, W, eng.pools[k]
            st = stacks[run_of[k][0]]
            i = k - run_of[k][0]
            blk.a = st["XA"][2 * i + 1]
            # LeakyReLU' masks of a (conv1 activation) and b (conv2 activation after Dropout2d) as sign bits
            blk.ma = torch.empty((B, H, W, F // 32), dtype=torch.int32, device=device) if train else None
            blk.mb = torch.empty((B, H, W, F // 32), dtype=torch.int32, device=device) if train else None
            blk.amax = None
            if blk.pool:
                blk.s = bf(H, W)
                H, W = H // 2, W // 2
                blk.out = out_buffer(k, H, W)
                if train:           # 2-bit positions of the window maxima: the un-pool does not re-read blk.s
                    blk.amax = torch.empty((B, H, W, F // 8), dtype=torch.int16, device=device)
            else:
                blk.s = out_buffer(k, H, W)
                blk.out = blk.s
            if train:
                blk.G = bf(H, W)
                blk.gs = bf(blk.H, blk.W) if blk.pool else None
                blk.gp1 = st["GP"][2 * i]
                blk.gp2 = st["GP"][2 * i + 1]
            self.blocks.append(blk)
        self.y = torch.empty((B, 5, eng.So_h, eng.So_w), dtype=F32, device=device)
        if train:
            # bf16 copy of the images in the stem's operand layout (written by the forward, read by the stem wgrad)
            n_cache = ops.stem_cache_elems(B, eng.in_ch, eng.in_h, eng.in_w, F, eng.stem_k, eng.stem_s, eng.stem_pad)
            self.x_cache = torch.zeros(n_cache, dtype=BF16, device=device) if n_cache else None
            self.g_stem = bf(eng.H0, eng.W0)
            self.loss = torch.empty((B,), dtype=F32, device=device)
            self.dy = torch.empty_like(self.y)
        self.generation = 0    # bumped by every forward that overwrites this plan (stale-backward detection)
        self.drop = None       # [num_blocks+1, B, F] fp32 Dropout2d multipliers (train mode only)
        self.x = None          # input of the last forward (needed by the stem wgrad)


class BackboneEngine:
    def __init__(self, filters: int, in_ch: int, in_h: int, in_w: int, num_blocks: int, stem_k: int, stem_s: int,
                 stem_pad: int, head_k: int, head_pad: int, pool_rule: Callable[[int], bool], slope: float = 0.2,
                 block_drop: float = 0.25, head_drop: float = 0.5):
        if filters != 64:
            raise NotImplementedError("the tcgen05 3x3 kernels are instantiated for 64 channels "
                                      "(the reference's 'medium' checkpoints); got filters=%d" % filters)
        self.F, self.in_ch, self.in_h, self.in_w = filters, in_ch, in_h, in_w
        self.F_logical = filters
        self.num_blocks, self.slope = num_blocks, slope
        self.stem_k, self.stem_s, self.stem_pad = stem_k, stem_s, stem_pad
        self.head_k, self.head_pad = head_k, head_pad
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
        self.So_h = H + 2 * head_pad - head_k + 1
        self.So_w = W + 2 * head_pad - head_k + 1
        # flat parameter layout (fp32)
        F = filters
        # w3 first: the gradient buffer is followed by the packed 3x3 accumulators, so [small sections | accumulators of
        # the layers before the fused chain] and [accumulators of the chain] are two CONTIGUOUS all-reduce regions
        self.sections = [
            ("w3", (2 * num_blocks, F, F, 3, 3)),
            ("conv1.weight", (F, in_ch, stem_k, stem_k)), ("conv1.bias", (F,)), ("b3", (2 * num_blocks, F)),
            ("out.weight", (5, F, head_k, head_k)), ("out.bias", (5,)),
        ]
        self.offsets, off = {}, 0
        for name, shape in self.sections:
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = (off, n, shape)
            off += (n + 3) // 4 * 4          # keep every section 16-byte aligned
        self.n_flat = off
        self.use_chain = True        # fuse runs of equal-shape blocks into one kernel when the image fits in smem
        # MaxPool2d(2) of a pooled block in the conv2 epilogue (fd_conv3x3_pool, bit-identical).  Measured on B200: the step
        # is 9 us SLOWER with it (502.8 vs 493.6 us) -- the extra barriers sit on the epilogue, which paces the MMA-bound
        # kernel, while the separate pool kernels are HBM-bound and cost 10 us -- so it is off by default.
        self.fuse_pool = False
        self.conv_flags = 0          # ops.CONV_ONE_TAP: per-layer launches bit-identical to the chain kernels (tests)
        self.device = None
        self.pflat = self.gflat = self.dwp = self.w_fwd = self.w_dgrad = None
        self.plans: Dict[tuple, _Plan] = {}
        self.weights_dirty = True

    # ------------------------------------------------------------------ parameters
    def param_names(self) -> List[str]:
        names = ["conv1.weight", "conv1.bias"]
        for k in range(self.num_blocks):
            for c in ("conv1", "conv2"):
                names += [f"residual_blocks.{k}.{c}.weight", f"residual_blocks.{k}.{c}.bias"]
        return names + ["out.weight", "out.bias"]

    def _view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        if name.startswith("residual_blocks."):
            _, k, c, kind = name.split(".")
            layer = 2 * int(k) + (0 if c == "conv1" else 1)
            off, n, shape = self.offsets["w3" if kind == "weight" else "b3"]
            return flat[off:off + n].view(shape)[layer]
        off, n, shape = self.offsets[name]
        return flat[off:off + n].view(shape)

    def section(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        off, n, shape = self.offsets[name]
        return flat[off:off + n].view(shape)

    def _ensure_device(self, device):
        if self.device == device and self.pflat is not None:
            return
        self.device = device
        self.pflat = torch.zeros(self.n_flat, dtype=F32, device=device)
        n3 = 2 * self.num_blocks * 9 * self.F * self.F
        # gradient buffer and the packed 3x3 weight-gradient accumulators share one allocation: one fill per step
        self.gzero = torch.zeros(self.n_flat + n3, dtype=F32, device=device)
        self.gflat = self.gzero[:self.n_flat]
        self.dwp = self.gzero[self.n_flat:]
        self.w_head_t = torch.empty(self.head_k * self.head_k * 5 * self.F, dtype=F32, device=device)
        self.w_fwd = torch.empty(n3, dtype=BF16, device=device)
        self.w_dgrad = torch.empty(n3, dtype=BF16, device=device)
        self.ar_stream = torch.cuda.Stream(device=device)        # the early gradient exchange runs beside the backward pass
        self.plans.clear()

    def bind(self, params: Dict[str, torch.nn.Parameter]):
        """Make every nn.Parameter a view of the flat buffer (values are preserved).  Cheap when
        nothing moved; re-flattens after ``.cuda()`` / ``.to()`` replaced the parameter storages."""
        dev = params["conv1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("the fd_b200 backbone runs on CUDA only (no CPU fallback): call model.cuda()")
        self._ensure_device(dev)
        for name in self.param_names():
            p = params[name]
            v = self._view(self.pflat, name)
            if p.data_ptr() != v.data_ptr():
                with torch.no_grad():
                    v.copy_(p.data.to(device=dev, dtype=F32))
                p.data = v
                self.weights_dirty = True

    def grad_view(self, name: str) -> torch.Tensor:
        return self._view(self.gflat, name)

    # ------------------------------------------------------------------ plans
    def plan(self, B: int, train: bool) -> _Plan:
        key = (B, train)
        if key not in self.plans:
            self.plans[key] = _Plan(self, B, train, self.device)
        return self.plans[key]

    def pack_weights(self):
        L = 2 * self.num_blocks
        ops.pack_conv3x3(self.section(self.pflat, "w3"), self.w_fwd.view(L, 9, self.F, self.F),
                         self.w_dgrad.view(L, 9, self.F, self.F))
        ops.head_pack(self.section(self.pflat, "out.weight"), self.w_head_t)
        self.weights_dirty = False

    def _wf(self, layer):
        n = 9 * self.F * self.F
        return self.w_fwd[layer * n:(layer + 1) * n]

    def _wd(self, layer):
        n = 9 * self.F * self.F
        return self.w_dgrad[layer * n:(layer + 1) * n]

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, train: bool, dropout: bool = False, repack: bool = True) -> _Plan:
        """x: [B,in_ch,H,W] fp32 in [0,1] (or uint8: /255 fused).  Returns the plan; plan.y is the
        sigmoid head [B,5,So,So] fp32.  ``dropout`` draws Dropout2d masks (train mode of the reference)."""
        B = x.shape[0]
        pl = self.plan(B, train)
        pl.generation += 1
        if repack or self.weights_dirty:
            self.pack_weights()
        if dropout:
            # one torch.rand (the caller's generator decides the masks) + one kernel for the multipliers
            r = torch.rand((self.num_blocks + 1, B, self.F), device=x.device)
            scale = torch.empty_like(r)
            ops.dropout_scale(r, self.num_blocks * B * self.F, 1.0 - self.block_drop, 1.0 - self.head_drop, scale)
            pl.drop = scale
        else:
            pl.drop = None
        self.run_forward(pl, x)
        return pl

    def run_forward(self, pl: _Plan, x: torch.Tensor):
        pl.x = x
        sb3 = self.section(self.pflat, "b3")
        ops.stem_fwd(x, self.section(self.pflat, "conv1.weight"), self.section(self.pflat, "conv1.bias"), pl.act0,
                     self.stem_s, self.stem_pad, x_cache=getattr(pl, "x_cache", None))
        cur = pl.act0
        chain_end = -1
        for k, blk in enumerate(pl.blocks):
            if k <= chain_end:
                cur = blk.out
                continue
            if k in pl.chains:
                chain_end = pl.chains[k]["k1"]
                self._chain_forward(pl, pl.chains[k], cur)
                cur = blk.out
                continue
            cs = pl.drop[k] if pl.drop is not None else None
            ops.conv3x3(cur, self._wf(2 * k), bias=sb3[2 * k], slope=self.slope, lrelu=True, mask_out=blk.ma,
                        out=blk.a, flags=self.conv_flags)
            if blk.pool and self.fuse_pool and blk.H % 2 == 0 and blk.W % 2 == 0:
                # MaxPool2d fused into the conv2 epilogue: the un-pooled sum blk.s is never written (the backward pass
                # un-pools from the recorded window positions)
                ops.conv3x3_pool(blk.a, self._wf(2 * k + 1), blk.out, bias=sb3[2 * k + 1], slope=self.slope,
                                 chan_scale=cs, residual=cur, mask_out=blk.mb, argmax=blk.amax)
            else:
                ops.conv3x3(blk.a, self._wf(2 * k + 1), bias=sb3[2 * k + 1], slope=self.slope, lrelu=True,
                            chan_scale=cs, residual=cur, mask_out=blk.mb, out=blk.s, flags=self.conv_flags)
                if blk.pool:
                    ops.maxpool2x2_fwd(blk.s, blk.out, blk.amax)
            cur = blk.out
        cs = pl.drop[self.num_blocks] if pl.drop is not None else None
        ops.head_fwd(cur, cs, self.section(self.pflat, "out.weight"), self.section(self.pflat, "out.bias"), pl.y,
                     self.head_pad, w_t=self.w_head_t)

    def _chain_forward(self, pl: _Plan, ch, x):
        k0, k1 = ch["k0"], ch["k1"]
        sb3 = self.section(self.pflat, "b3")
        descs = []
        for k in range(k0, k1 + 1):
            blk = pl.blocks[k]
            d = {"bias1": sb3[2 * k], "bias2": sb3[2 * k + 1],
                 "chan_scale": pl.drop[k] if pl.drop is not None else None}
            if pl.train:
                d.update(a=blk.a, mask_a=blk.ma, mask_b=blk.mb, out=blk.s)
            elif k == k1:
                d.update(out=blk.s)
            descs.append(d)
        n3 = 9 * self.F * self.F
        ops.resblock_chain_fwd(x, self.w_fwd[2 * k0 * n3:(2 * k1 + 2) * n3], descs, self.slope)

    def _chain_backward(self, pl: _Plan, ch, g_in):
        """Input-gradient chain of blocks k1..k0 in one launch; G / gp2 of block k1 are already in place."""
        k0, k1 = ch["k0"], ch["k1"]
        drop = pl.drop
        descs = []
        for k in range(k0, k1 + 1):
            blk = pl.blocks[k]
            d = {"mask_a": blk.ma, "gp1": blk.gp1}
            if k == k0:
                d["g_in"] = g_in
            else:
                d.update(mask_b_prev=pl.blocks[k - 1].mb, gp2_prev=pl.blocks[k - 1].gp2,
                         chan_scale_prev=drop[k - 1] if drop is not None else None)
            descs.append(d)
        n3 = 9 * self.F * self.F
        last = pl.blocks[k1]
        ops.resblock_chain_bwd(last.G, last.gp2, self.w_dgrad[2 * k0 * n3:(2 * k1 + 2) * n3], descs, self.slope)

    # ------------------------------------------------------------------ backward
    def exchange_regions(self, pl: _Plan):
        """(early, late): the two contiguous slices of the gradient allocation a data-parallel step all-reduces.
        `early` = packed 3x3 weight-gradient accumulators of the fused chain (final as soon as the chain's weight-gradient
        launch returns, ~2/3 of the backward pass before its end; 77 % of the gradient bytes of PoolResnet-medium);
        `late` = every other gradient (stem, biases, head, accumulators of the layers in front of the chain).  The
        unpacked w3 section of gflat is NOT exchanged: it is produced from the reduced accumulators afterwards."""
        n3 = 9 * self.F * self.F
        w3_end = self.offsets["w3"][0] + self.offsets["w3"][1]
        split = self.n_flat + self.dwp.numel()
        if pl.chains:
            k0 = min(ch["k0"] for ch in pl.chains.values())
            k1 = max(ch["k1"] for ch in pl.chains.values())
            if k1 == self.num_blocks - 1:
                split = self.n_flat + 2 * k0 * n3
        return self.gzero[split:], self.gzero[w3_end:split]

    def run_backward(self, pl: _Plan, dy: torch.Tensor, exchange=None):
        """dy: gradient w.r.t. plan.y.  Fills self.gflat (overwrites).  ``exchange`` (parallel.SplitAllReduce): the
        data-parallel gradient sum, overlapped -- `exchange.early` runs on a side stream while the 30x30 / 60x60 layers
        are still being differentiated, `exchange.late` just before the accumulators are unpacked."""
        assert pl.train
        early_t, late_t = self.exchange_regions(pl) if exchange is not None else (None, None)
        early_done = exchange is None or early_t.numel() == 0
        nb = self.num_blocks
        self.gzero.zero_()
        gb3 = self.section(self.gflat, "b3")
        n3 = 9 * self.F * self.F
        drop = pl.drop
        last = pl.blocks[nb - 1]
        ops.head_bwd(last.out, drop[nb] if drop is not None else None, self.section(self.pflat, "out.weight"), pl.y,
                     dy, self.head_pad, last.G, None if last.pool else last.mb,
                     None if (last.pool or drop is None) else drop[nb - 1], self.slope,
                     None if last.pool else last.gp2, self.section(self.gflat, "out.weight"),
                     self.section(self.gflat, "out.bias"), w_t=self.w_head_t)
        group_of_first = {grp[0]: grp for grp in pl.groups}
        gb3_flat = gb3.reshape(-1)

        def group_wgrad(k0):
            """All weight gradients of the run starting at block k0 (2 per block) in ONE multi-problem launch."""
            _, _, XA_all, GP_all = group_of_first[k0]
            ops.conv3x3_wgrad_multi(XA_all, GP_all, self.dwp[(2 * k0) * n3:], n3, gb3_flat[(2 * k0) * self.F:], self.F)

        chain_of_last = {ch["k1"]: ch for ch in pl.chains.values()}
        skip_until = nb
        for k in range(nb - 1, -1, -1):
            blk = pl.blocks[k]
            if k in chain_of_last and not blk.pool:
                ch = chain_of_last[k]
                k0 = ch["k0"]
                self._chain_backward(pl, ch, pl.blocks[k0 - 1].G if k0 > 0 else pl.g_stem)
                skip_until = k0
                group_wgrad(k0)
                if exchange is not None and not early_done and k == self.num_blocks - 1:
                    main = torch.cuda.current_stream()
                    self.ar_stream.wait_stream(main)
                    with torch.cuda.stream(self.ar_stream):
                        exchange.early(early_t)
                    early_done = True
                if k0 == 0:
                    ops.stem_wgrad(pl.x, pl.g_stem, self.section(self.gflat, "conv1.weight"),
                                   self.section(self.gflat, "conv1.bias"), self.stem_s, self.stem_pad,
                                   x_cache=getattr(pl, "x_cache", None))
                continue
            if k >= skip_until:
                continue
            if blk.pool:
                ops.maxpool2x2_bwd(blk.s, blk.G, blk.gs, blk.mb, drop[k] if drop is not None else None, self.slope,
                                   blk.gp2, argmax=blk.amax)
                GS = blk.gs
            else:
                GS = blk.G
            cf = self.conv_flags
            ops.conv3x3(blk.gp2, self._wd(2 * k + 1), slope=self.slope, mask_in=blk.ma, out2=blk.gp1, flags=cf)
            if k > 0:
                prev = pl.blocks[k - 1]
                if prev.pool:
                    ops.conv3x3(blk.gp1, self._wd(2 * k), slope=self.slope, residual=GS, out=prev.G, flags=cf)
                else:
                    ops.conv3x3(blk.gp1, self._wd(2 * k), slope=self.slope, residual=GS, out=prev.G,
                                mask_in=prev.mb, chan_scale2=drop[k - 1] if drop is not None else None,
                                out2=prev.gp2, flags=cf)
            else:
                ops.conv3x3(blk.gp1, self._wd(0), slope=self.slope, residual=GS, out=pl.g_stem, flags=cf)
                ops.stem_wgrad(pl.x, pl.g_stem, self.section(self.gflat, "conv1.weight"),
                               self.section(self.gflat, "conv1.bias"), self.stem_s, self.stem_pad,
                               x_cache=getattr(pl, "x_cache", None))
            if k in group_of_first:
                group_wgrad(k)          # gp1 / gp2 of every block of the run are final now
        if exchange is not None:
            if early_t.numel():
                torch.cuda.current_stream().wait_stream(self.ar_stream)
            exchange.late(late_t)
        ops.unpack_wgrad3x3(self.dwp.view(2 * nb, 9, self.F, self.F), self.section(self.gflat, "w3"))

    # ------------------------------------------------------------------ fused train step
    def train_step(self, x: torch.Tensor, gt: torch.Tensor, dropout: bool = True, allreduce=None,
                   optimizer=None) -> _Plan:
        """forward -> per-image YoloLoss (+ its gradient, same kernel) -> backward [-> gradient all-reduce over the
        data-parallel ranks -> optimizer step].  Afterwards plan.loss holds the per-image losses and self.gflat the
        gradient of their SUM (models/ModelMeta.py:173-176,215: the reference sums, it does not average).
        ``allreduce``: callable on the flat gradient buffer (parallel.PeerAllReduce or parallel.allreduce_grads);
        ``optimizer``: optim.FlatAdam (models/ModelMeta.py:104-112)."""
        with torch.cuda.device(x.device):
            return self._train_step(x, gt, dropout, allreduce, optimizer)

    def _train_step(self, x, gt, dropout, allreduce, optimizer) -> _Plan:
        pl = self.forward(x, train=True, dropout=dropout)
        if tuple(gt.shape) != tuple(pl.y.shape) or gt.device != pl.y.device:
            raise ValueError(f"target map {tuple(gt.shape)} on {gt.device} does not match the head "
                             f"{tuple(pl.y.shape)} on {pl.y.device}")
        ops.yolo_loss(pl.y, gt, pl.loss, None, pl.dy)
        if allreduce is not None and hasattr(allreduce, "early"):
            self.run_backward(pl, pl.dy, exchange=allreduce)       # overlapped exchange of the packed accumulators
        else:
            self.run_backward(pl, pl.dy)
            if allreduce is not None:
                allreduce(self.opt_grads())
        if optimizer is not None:
            optimizer.step()
        return pl

    # flat buffers the optimizer and the data-parallel all-reduce work on (PaddedBackboneEngine: the un-padded ones)
    def opt_params(self) -> torch.Tensor:
        return self.pflat

    def opt_grads(self) -> torch.Tensor:
        return self.gflat

    # ------------------------------------------------------------------ CUDA graph of the train step
    def capture_train_step(self, x_static: torch.Tensor, gt_static: torch.Tensor, dropout: bool = True,
                           allreduce=None, optimizer=None):
        """Capture forward + loss + backward (+ peer-memory all-reduce + Adam) on fixed input buffers into one CUDA
        graph (kills the launch gaps and all host-side descriptor encoding).  Returns (graph, plan,
        launches_per_step).  An optimizer inside a graph must keep its step count on the device
        (optim.FlatAdam(capturable=True))."""
        from .native import launch_count
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                self.train_step(x_static, gt_static, dropout, allreduce, optimizer)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        with torch.cuda.graph(graph):
            pl = self.train_step(x_static, gt_static, dropout, allreduce, optimizer)
        return graph, pl, launch_count() - n0


class PaddedBackboneEngine(BackboneEngine):
    """Backbones NARROWER than the 64-channel kernel planes (the reference's "small" checkpoint is filters = 32,
    saved_models/official/PoolResnet/small_model_10x10_480.pth) on the 64-channel engine: every weight tensor is
    embedded in a zero-padded 64-channel one.  Padded channels stay exactly zero through LeakyReLU / skip / pooling and
    their weight gradients are exactly zero, so the arithmetic of the logical channels is unchanged.  The
    ``nn.Parameter`` tensors are views of an un-padded flat buffer ``psmall`` (source of truth, optimizer / all-reduce
    unit); one ``fd_index_copy_f32`` scatters it into the padded buffer before a forward pass and one gathers the
    gradients back after a backward pass."""

    def __init__(self, filters: int, in_ch, in_h, in_w, num_blocks, stem_k, stem_s, stem_pad, head_k, head_pad, pool_rule,
                 **kw):
        if not (0 < filters < 64):
            raise NotImplementedError("PaddedBackboneEngine handles 0 < filters < 64")
        super().__init__(64, in_ch, in_h, in_w, num_blocks, stem_k, stem_s, stem_pad, head_k, head_pad, pool_rule, **kw)
        f = self.F_logical = filters
        self.small_sections = [
            ("conv1.weight", (f, in_ch, stem_k, stem_k)), ("conv1.bias", (f,)),
            ("w3", (2 * num_blocks, f, f, 3, 3)), ("b3", (2 * num_blocks, f)),
            ("out.weight", (5, f, head_k, head_k)), ("out.bias", (5,)),
        ]
        self.small_offsets, off = {}, 0
        for name, shape in self.small_sections:
            n = 1
            for d in shape:
                n *= d
            self.small_offsets[name] = (off, n, shape)
            off += (n + 3) // 4 * 4
        self.n_small = off
        self.psmall = self.gsmall = self.index = None

    def _small_view(self, flat, name):
        if name.startswith("residual_blocks."):
            _, k, c, kind = name.split(".")
            layer = 2 * int(k) + (0 if c == "conv1" else 1)
            off, n, shape = self.small_offsets["w3" if kind == "weight" else "b3"]
            return flat[off:off + n].view(shape)[layer]
        off, n, shape = self.small_offsets[name]
        return flat[off:off + n].view(shape)

    def _ensure_device(self, device):
        if self.device == device and self.pflat is not None and self.psmall is not None:
            return
        super()._ensure_device(device)
        f = self.F_logical
        self.psmall = torch.zeros(self.n_small, dtype=F32, device=device)
        self.gsmall = torch.zeros(self.n_small, dtype=F32, device=device)
        # index[i] = position of small element i in the padded flat buffer (alignment gaps map onto themselves' twin:
        # they point at padded zeros of the same section and carry zeros)
        big = torch.arange(self.n_flat, dtype=torch.int32, device=device)
        idx = torch.zeros(self.n_small, dtype=torch.int32, device=device)
        sel = {"conv1.weight": lambda v: v[:f], "conv1.bias": lambda v: v[:f], "w3": lambda v: v[:, :f, :f],
               "b3": lambda v: v[:, :f], "out.weight": lambda v: v[:, :f], "out.bias": lambda v: v}
        for name, _ in self.small_sections:
            off, n, shape = self.small_offsets[name]
            boff, bn, bshape = self.offsets[name]
            src = sel[name](big[boff:boff + bn].view(bshape)).reshape(-1)
            idx[off:off + n] = src
            gap = (n + 3) // 4 * 4 - n
            if gap:       # alignment gap of the small buffer (holds zeros): aim at the padded buffer's own gap, or at the
                          # section's last element, which belongs to a padded (zero) channel
                big_gap = (bn + 3) // 4 * 4 - bn
                for j in range(gap):
                    idx[off + n + j] = boff + bn + j if j < big_gap else boff + bn - 1
                assert big_gap >= gap or name != "out.bias"
        self.index = idx.contiguous()

    def bind(self, params):
        dev = params["conv1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("the fd_b200 backbone runs on CUDA only (no CPU fallback): call model.cuda()")
        self._ensure_device(dev)
        for name in self.param_names():
            p = params[name]
            v = self._small_view(self.psmall, name)
            if p.data_ptr() != v.data_ptr():
                with torch.no_grad():
                    v.copy_(p.data.to(device=dev, dtype=F32))
                p.data = v
        self.weights_dirty = True

    def grad_view(self, name):
        return self._small_view(self.gsmall, name)

    def opt_params(self):
        return self.psmall

    def opt_grads(self):
        return self.gsmall

    def pack_weights(self):
        ops.index_copy(self.pflat, self.psmall, self.index, scatter=True)     # padded positions stay zero
        super().pack_weights()

    def forward(self, x, train, dropout=False, repack=True):
        return super().forward(x, train, dropout=dropout, repack=True)

    def run_backward(self, pl, dy, exchange=None):
        super().run_backward(pl, dy, exchange=exchange)
        ops.index_copy(self.gsmall, self.gflat, self.index, scatter=False)
