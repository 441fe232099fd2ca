"""Execution engine of the SSD model (reference models/SSD.py:14-255): stem conv, nine + four ``SeparableResidualBlock``s
with channel counts 16 ... 256 (1x1 skip convolution where in != out), four ``Linear(C -> 5)`` heads with
``apply_priors`` -- forward, backward and the train step of ``ModelMetaSSD.step`` (models/ModelMetaSSD.py:175).

Every activation is stored as G = ceil(C / 64) NHWC bf16 PLANES of 64 channels (channel counts below 64 are zero padded:
padded channels stay exactly zero through LeakyReLU / skip / pooling and their weight gradients are exactly zero), so the
tcgen05 kernels instantiated for 64 channels are reused unchanged:

* a Cin -> Cout 3x3 convolution = per output plane g the sum over input planes h of 64 -> 64 convolutions with the
  (zero-padded) weight sub-block W[g][h]; the sum is chained through ``fd_conv3x3``'s residual operand.  With ONE input
  plane (the nine feature-extractor blocks, i.e. all 240 / 120 / 60-pixel-wide layers below 128 channels) bias, LeakyReLU,
  Dropout2d multiplier, sign-bit mask and skip add are fused into the convolution's epilogue; otherwise they run in
  ``fd_act_mask`` / ``fd_grad_mask``;
* the 1x1 skip convolutions are ``fd_conv3x3`` in centre-tap mode (``FD_CONV_1X1``), their weight gradients the centre
  tap of ``fd_conv3x3_wgrad``;
* layers with at least 128 output channels (the four ``continue_layers`` blocks: 64 -> 128, 128 -> 256, 256 -> 256 -- 80 %
  of the model's FLOPs) run on ``fd_conv3x3_wide`` instead: one launch per group of 128 output channels
  (tcgen05.mma.cta_group::2, N = 128, the sum over input planes in TMEM, epilogue fused), packed straight from the
  un-padded parameter views (``use_wide``); their weight gradients stay on the 64-channel ``fd_conv3x3_wgrad``;
* parameters live un-padded in ONE flat fp32 buffer ``pflat`` (the optimizer / data-parallel all-reduce unit; every
  ``nn.Parameter`` is a view), scattered into the zero-padded packing buffer by one ``fd_index_copy_f32``; gradients are
  gathered back by one ``fd_index_copy_f32`` into ``gflat`` (every ``p.grad`` is a view).
"""
from __future__ import annotations

from typing import Dict, List

import contextlib
import os

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


def _planes_of(c):
    return (c + 63) // 64


class _SBlock:
    """Static description of one SeparableResidualBlock (models/SSD.py:14-81)."""

    def __init__(self, name, cin, cout, pool):
        self.name, self.cin, self.cout, self.pool = name, cin, cout, pool
        self.gi, self.go = _planes_of(cin), _planes_of(cout)
        self.has_skip_conv = cin != cout


class SSDEngine:
    def __init__(self, filters: int, in_ch: int, in_h: int, in_w: int, slope: float = 0.2, block_drop: float = 0.25):
        f = filters
        self.f, self.in_ch, self.in_h, self.in_w, self.slope, self.block_drop = f, in_ch, in_h, in_w, slope, block_drop
        mx = 16 * f
        blocks = [_SBlock("feature_extractor.0", f, 2 * f, True), _SBlock("feature_extractor.1", 2 * f, 2 * f, True)]
        blocks += [_SBlock(f"feature_extractor.{i}", 2 * f, 2 * f, False) for i in range(2, 8)]
        blocks += [_SBlock("feature_extractor.8", 2 * f, 4 * f, False)]
        self.head_channels = []
        for i in range(4):
            cin = min(4 * f * (2 ** i), mx)
            cout = min(2 * cin, mx)
            blocks.append(_SBlock(f"continue_layers.{i}.0", cin, cout, i != 0))
            self.head_channels.append(cout)
        for b in blocks:
            for c in (b.cin, b.cout):
                if c > 64 and c % 64:
                    raise NotImplementedError(f"SSD channel count {c} is neither <= 64 nor a multiple of 64")
        if f > 64:
            raise NotImplementedError("SSD stem wider than 64 channels")
        self.blocks = blocks
        self.n_fe = 9
        self.H0, self.W0 = (in_h + 2 - 3) // 2 + 1, (in_w + 2 - 3) // 2 + 1
        # spatial sizes
        h, w = self.H0, self.W0
        self.shapes = []
        for b in blocks:
            self.shapes.append((h, w))
            if b.pool:
                h, w = h // 2, w // 2
        self.head_hw = []
        h, w = self.H0, self.W0
        for i, b in enumerate(blocks):
            if b.pool:
                h, w = h // 2, w // 2
            if i >= self.n_fe:
                self.head_hw.append((h, w))
        self.patch_sizes = tuple(hw[0] for hw in self.head_hw)
        self.P = sum(hh * ww for hh, ww in self.head_hw)
        # ---- un-padded flat parameter layout (names = the reference's state_dict keys)
        self.sections, off = {}, 0

        def add(name, shape):
            nonlocal off
            n = 1
            for d in shape:
                n *= d
            self.sections[name] = (off, n, tuple(shape))
            off += (n + 3) // 4 * 4

        add("input_normalizer.weight", (f, in_ch, 3, 3))
        add("input_normalizer.bias", (f,))
        for b in blocks:
            if b.has_skip_conv:
                add(b.name + ".pointwise_conv_skip.weight", (b.cout, b.cin, 1, 1))
                add(b.name + ".pointwise_conv_skip.bias", (b.cout,))
            add(b.name + ".conv1.weight", (b.cout, b.cin, 3, 3))
            add(b.name + ".conv1.bias", (b.cout,))
            add(b.name + ".conv2.weight", (b.cout, b.cout, 3, 3))
            add(b.name + ".conv2.bias", (b.cout,))
        for i, c in enumerate(self.head_channels):
            add(f"extracting_layers.{i}.0.weight", (5, c))
            add(f"extracting_layers.{i}.0.bias", (5,))
        self.n_flat = off
        # ---- padded packing layout: stem [64,in_ch,3,3] + [64]; per conv layer: sub-blocks [(g,h),64,64,3,3] + bias [go,64]
        self.conv_layers = []          # (param prefix, cout, cin, go, gi, first sub-block, is_1x1)
        nsub = 0
        for b in blocks:
            if b.has_skip_conv:
                self.conv_layers.append((b.name + ".pointwise_conv_skip", b.cout, b.cin, b.go, b.gi, nsub, True))
                nsub += b.go * b.gi
            self.conv_layers.append((b.name + ".conv1", b.cout, b.cin, b.go, b.gi, nsub, False))
            nsub += b.go * b.gi
            self.conv_layers.append((b.name + ".conv2", b.cout, b.cout, b.go, b.go, nsub, False))
            nsub += b.go * b.go
        self.n_sub = nsub
        self.layer_of = {l[0]: l for l in self.conv_layers}
        self.pad_w3_off = 0
        self.pad_b3_off = nsub * 64 * 64 * 9
        self.bias_row = {}
        rows = 0
        for l in self.conv_layers:
            self.bias_row[l[0]] = rows
            rows += l[3]
        self.n_bias_rows = rows
        self.pad_stem_w_off = self.pad_b3_off + rows * 64
        self.pad_stem_b_off = self.pad_stem_w_off + 64 * in_ch * 9
        self.n_pad = self.pad_stem_b_off + 64 + 4      # + a spare zero the alignment gaps of the un-padded buffer map to
        self.device = None
        self.pflat = self.gflat = None
        self.use_wide = True           # fd_conv3x3_wide for layers whose output (forward) / input (dgrad) is 128k channels
        self.plans: Dict[tuple, dict] = {}

    # ------------------------------------------------------------------ parameters
    def param_names(self) -> List[str]:
        return list(self.sections.keys())

    def _view(self, flat, name):
        off, n, shape = self.sections[name]
        return flat[off:off + n].view(shape)

    def grad_view(self, name):
        return self._view(self.gflat, name)

    def _build_index(self, device):
        """index[i] = position of un-padded element i in the padded packing buffer (built with tensor views, once)."""
        big = torch.arange(self.n_pad, dtype=torch.int32, device=device)
        idx = torch.full((self.n_flat,), self.n_pad - 1, dtype=torch.int32, device=device)   # gaps -> a padded zero
        f = self.f

        def put(name, src):
            off, n, _ = self.sections[name]
            idx[off:off + n] = src.reshape(-1)

        put("input_normalizer.weight", big[self.pad_stem_w_off:self.pad_stem_w_off + 64 * self.in_ch * 9]
            .view(64, self.in_ch, 3, 3)[:f])
        put("input_normalizer.bias", big[self.pad_stem_b_off:self.pad_stem_b_off + 64][:f])
        for (pre, cout, cin, go, gi, first, is1) in self.conv_layers:
            sub = big[first * 36864:(first + go * gi) * 36864].view(go, gi, 64, 64, 3, 3)
            # logical [cout, cin, k, k] -> (g, co) x (h, ci): permute the padded view to [go, 64, gi, 64, 3, 3]
            full = sub.permute(0, 2, 1, 3, 4, 5).reshape(go * 64, gi * 64, 3, 3)[:cout, :cin]
            put(pre + ".weight", full[:, :, 1:2, 1:2] if is1 else full)
            r0 = self.bias_row[pre]
            put(pre + ".bias", big[self.pad_b3_off + r0 * 64:self.pad_b3_off + (r0 + go) * 64][:cout])
        self.index = idx.contiguous()
        self.conv_mask = torch.ones(self.n_flat, dtype=torch.bool, device=device)
        for i in range(4):                       # head parameters are used un-padded, straight from pflat
            for k in ("weight", "bias"):
                off, n, _ = self.sections[f"extracting_layers.{i}.0.{k}"]
                self.conv_mask[off:off + n] = False
        # compact index lists for the conv parameters only
        self.conv_pos = torch.nonzero(self.conv_mask).flatten().to(torch.int32)
        self.index_conv = self.index[self.conv_pos.long()].contiguous()

    def bind(self, params):
        dev = params["input_normalizer.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("the fd_b200 SSD model runs on CUDA only (no CPU fallback): call model.cuda()")
        if self.device != dev or self.pflat is None:
            self.device = dev
            self.pflat = torch.zeros(self.n_flat, dtype=F32, device=dev)
            self.gflat = torch.zeros(self.n_flat, dtype=F32, device=dev)
            self.ppad = torch.zeros(self.n_pad, dtype=F32, device=dev)
            self.gpad = torch.zeros(self.n_pad, dtype=F32, device=dev)
            self.psrc = None
            self.dwp = torch.zeros((self.n_sub, 9 * 64 * 64), dtype=F32, device=dev)
            self.w_fwd = torch.empty((self.n_sub, 9, 64, 64), dtype=BF16, device=dev)
            self.w_dgrad = torch.empty((self.n_sub, 9, 64, 64), dtype=BF16, device=dev)
            # wide packings [group of 128][input plane][tap][128][64] of the layers the wide kernel serves
            self.wide_f, self.wide_d = {}, {}
            for (pre, cout, cin, go, gi, first, is1) in self.conv_layers:
                if go % 2 == 0 and gi <= 4:
                    self.wide_f[pre] = torch.zeros((go // 2, gi, 9, 128, 64), dtype=BF16, device=dev)
                if gi % 2 == 0 and go <= 4:
                    self.wide_d[pre] = torch.zeros((gi // 2, go, 9, 128, 64), dtype=BF16, device=dev)
            self._build_index(dev)
            self.pconv = torch.zeros(self.conv_pos.numel(), dtype=F32, device=dev)
            self.gconv = torch.zeros(self.conv_pos.numel(), dtype=F32, device=dev)
            self.plans.clear()
        for name in self.param_names():
            p = params[name]
            v = self._view(self.pflat, name)
            if p.data_ptr() != v.data_ptr():
                with torch.no_grad():
                    v.copy_(p.data.to(device=dev, dtype=F32))
                p.data = v

    def opt_params(self):
        return self.pflat

    def opt_grads(self):
        return self.gflat

    def pack_weights(self):
        # un-padded flat -> compact conv list (gather) -> padded packing buffer (scatter) -> bf16 operand layouts
        ops.index_copy(self.pconv, self.pflat, self.conv_pos, scatter=False)
        ops.index_copy(self.ppad, self.pconv, self.index_conv, scatter=True)
        ops.pack_conv3x3(self.ppad[:self.pad_b3_off].view(self.n_sub, 64, 64, 3, 3), self.w_fwd, self.w_dgrad)
        if self.use_wide:
            for (pre, cout, cin, go, gi, first, is1) in self.conv_layers:
                if pre in self.wide_f or pre in self.wide_d:
                    ops.pack_conv3x3_wide(self._view(self.pflat, pre + ".weight"), self.wide_f.get(pre), self.wide_d.get(pre))

    def _bias2(self, pre, gg):
        """bias of output planes 2gg, 2gg+1 (128 contiguous floats of the padded buffer)"""
        r = self.bias_row[pre] + 2 * gg
        return self.ppad[self.pad_b3_off + r * 64:self.pad_b3_off + (r + 2) * 64]

    @staticmethod
    def _pair(planes, gg):
        return None if planes is None or planes[0] is None else [planes[2 * gg], planes[2 * gg + 1]]

    def _bias(self, pre, g):
        r = self.bias_row[pre] + g
        return self.ppad[self.pad_b3_off + r * 64:self.pad_b3_off + (r + 1) * 64]

    def _gbias(self, pre, g):
        r = self.bias_row[pre] + g
        return self.gpad[self.pad_b3_off + r * 64:self.pad_b3_off + (r + 1) * 64]

    def _sub(self, pre, g, h):
        l = self.layer_of[pre]
        return l[5] + g * l[4] + h

    # ------------------------------------------------------------------ plans
    def plan(self, B, train):
        key = (B, train)
        if key in self.plans:
            return self.plans[key]
        dev = self.device

        def planes(n, h, w):
            return [torch.empty((B, h, w, 64), dtype=BF16, device=dev) for _ in range(n)]

        def masks(n, h, w):
            return [torch.empty((B, h, w, 2), dtype=torch.int32, device=dev) for _ in range(n)]

        pl = {"B": B, "train": train, "generation": 0, "blocks": []}
        pl["act0"] = planes(1, self.H0, self.W0)
        for i, b in enumerate(self.blocks):
            h, w = self.shapes[i]
            d = {"T": planes(b.go, h, w), "T2": planes(b.go, h, w), "a": planes(b.go, h, w), "s": planes(b.go, h, w)}
            d["skip"] = planes(b.go, h, w) if b.has_skip_conv else None
            d["out"] = planes(b.go, h // 2, w // 2) if b.pool else d["s"]
            if train:
                d["ma"], d["mb"] = masks(b.go, h, w), masks(b.go, h, w)
                d["amax"] = [torch.empty((B, h // 2, w // 2, 8), dtype=torch.int16, device=dev) for _ in range(b.go)] if b.pool else None
                ho, wo = (h // 2, w // 2) if b.pool else (h, w)
                d["G"] = planes(b.go, ho, wo)
                d["gs"] = planes(b.go, h, w) if b.pool else None
                d["gp1"], d["gp2"] = planes(b.go, h, w), planes(b.go, h, w)
                d["U"] = planes(max(b.gi, b.go), h, w)
                d["U2"] = planes(max(b.gi, b.go), h, w)
            pl["blocks"].append(d)
        pl["y"] = torch.empty((B, self.P, 5), dtype=F32, device=dev)
        if train:
            pl["g_stem"] = planes(1, self.H0, self.W0)
            pl["head_dx"] = [planes(_planes_of(c), hh, ww) for c, (hh, ww) in zip(self.head_channels, self.head_hw)]
            pl["dy"] = torch.empty_like(pl["y"])
            pl["sums"] = torch.empty((B, 2), dtype=F32, device=dev)
            pl["npos"] = torch.empty((B,), dtype=torch.int32, device=dev)
            pl["dconf"] = torch.empty((B, self.P), dtype=F32, device=dev)
            pl["dloc"] = torch.empty((B, self.P, 4), dtype=F32, device=dev)
            pl["loss"] = torch.empty((), dtype=F32, device=dev)
        pl["drop"] = None
        self.plans[key] = pl
        return pl

    # ------------------------------------------------------------------ forward
    def _conv_sum(self, srcs, wsel, pre, g, gi, bias, dst_a, dst_b, flags=0, first_residual=None, **last_kw):
        """sum_h conv(srcs[h], W[pre][g][h]) (+ bias) (+ first_residual) as a chain of 64-channel convolutions; the last
        call takes ``last_kw`` (fused epilogue).  Returns the buffer holding the result."""
        prev = first_residual
        dst = None
        for h in range(gi):
            last = h == gi - 1
            dst = dst_b if (h % 2) else dst_a
            kw = dict(bias=bias if h == 0 else None, slope=self.slope, lrelu=False, residual=prev, flags=flags)
            if last and last_kw:
                kw.update(last_kw)
                if "out" not in kw and "out2" not in kw:
                    kw["out"] = dst
            else:
                kw["out"] = dst
            ops.conv3x3(srcs[h], wsel[self._sub(pre, g, h)], **kw)
            prev = kw.get("out", dst)
        return prev

    def forward(self, x, train: bool, dropout: bool = False, priors=None, mult=None):
        B = x.shape[0]
        pl = self.plan(B, train)
        pl["generation"] += 1
        self.pack_weights()
        if dropout:
            tot = sum(b.go for b in self.blocks)
            r = torch.rand((tot, B, 64), device=x.device)
            scale = torch.empty_like(r)
            ops.dropout_scale(r, r.numel(), 1.0 - self.block_drop, 1.0, scale)
            pl["drop"], k = [], 0
            for b in self.blocks:
                pl["drop"].append([scale[k + g] for g in range(b.go)])
                k += b.go
        else:
            pl["drop"] = None
        pl["x"] = x
        sw = self.ppad[self.pad_stem_w_off:self.pad_stem_b_off].view(64, self.in_ch, 3, 3)
        sb = self.ppad[self.pad_stem_b_off:self.pad_stem_b_off + 64]
        ops.stem_fwd(x, sw, sb, pl["act0"][0], 2, 1)
        cur = pl["act0"]
        hi = 0
        off = 0
        for i, b in enumerate(self.blocks):
            d = pl["blocks"][i]
            d["inp"] = cur
            drop = pl["drop"][i] if pl["drop"] is not None else [None] * b.go
            ma = d.get("ma") or [None] * b.go
            mb = d.get("mb") or [None] * b.go
            pre1, pre2, pres = b.name + ".conv1", b.name + ".conv2", b.name + ".pointwise_conv_skip"
            if self.use_wide and pre1 in self.wide_f and pre2 in self.wide_f:
                for gg in range(b.go // 2):
                    if b.has_skip_conv:
                        ops.conv3x3_wide(cur, self.wide_f[pres][gg], bias=self._bias2(pres, gg), slope=self.slope,
                                         out=self._pair(d["skip"], gg), flags=ops.CONV_1X1)
                    ops.conv3x3_wide(cur, self.wide_f[pre1][gg], bias=self._bias2(pre1, gg), slope=self.slope, lrelu=True,
                                     mask_out=self._pair(ma, gg), out=self._pair(d["a"], gg))
                skip = d["skip"] if b.has_skip_conv else cur
                for gg in range(b.go // 2):
                    ops.conv3x3_wide(d["a"], self.wide_f[pre2][gg], bias=self._bias2(pre2, gg), slope=self.slope, lrelu=True,
                                     chan_scale=self._pair(drop, gg), residual=self._pair(skip, gg),
                                     mask_out=self._pair(mb, gg), out=self._pair(d["s"], gg))
                if b.pool:
                    for g in range(b.go):
                        ops.maxpool2x2_fwd(d["s"][g], d["out"][g], d["amax"][g] if train else None)
                cur = d["out"]
                if i >= self.n_fe:
                    w = self._view(self.pflat, f"extracting_layers.{hi}.0.weight")
                    bb = self._view(self.pflat, f"extracting_layers.{hi}.0.bias")
                    ops.ssd_head_fwd(cur, w, bb, mult, priors, off, pl["y"])
                    off += self.head_hw[hi][0] * self.head_hw[hi][1]
                    hi += 1
                continue
            if b.has_skip_conv:
                for g in range(b.go):
                    self._conv_sum(cur, self.w_fwd, b.name + ".pointwise_conv_skip", g, b.gi,
                                   self._bias(b.name + ".pointwise_conv_skip", g), d["T"][g], d["T2"][g],
                                   flags=ops.CONV_1X1, out=d["skip"][g])
                skip = d["skip"]
            else:
                skip = cur
            for g in range(b.go):
                bias = self._bias(b.name + ".conv1", g)
                if b.gi == 1:
                    ops.conv3x3(cur[0], self.w_fwd[self._sub(b.name + ".conv1", g, 0)], bias=bias, slope=self.slope,
                                lrelu=True, mask_out=ma[g], out=d["a"][g])
                else:
                    raw = self._conv_sum(cur, self.w_fwd, b.name + ".conv1", g, b.gi, bias, d["T"][g], d["T2"][g])
                    ops.act_mask(raw, self.slope, None, None, ma[g], d["a"][g])
            for g in range(b.go):
                bias = self._bias(b.name + ".conv2", g)
                if b.go == 1:
                    ops.conv3x3(d["a"][0], self.w_fwd[self._sub(b.name + ".conv2", g, 0)], bias=bias, slope=self.slope,
                                lrelu=True, chan_scale=drop[g], residual=skip[g], mask_out=mb[g], out=d["s"][g])
                else:
                    raw = self._conv_sum(d["a"], self.w_fwd, b.name + ".conv2", g, b.go, bias, d["T"][g], d["T2"][g])
                    ops.act_mask(raw, self.slope, drop[g], skip[g], mb[g], d["s"][g])
                if b.pool:
                    ops.maxpool2x2_fwd(d["s"][g], d["out"][g], d["amax"][g] if train else None)
            cur = d["out"]
            if i >= self.n_fe:
                w = self._view(self.pflat, f"extracting_layers.{hi}.0.weight")
                bb = self._view(self.pflat, f"extracting_layers.{hi}.0.bias")
                ops.ssd_head_fwd(cur, w, bb, mult, priors, off, pl["y"])
                off += self.head_hw[hi][0] * self.head_hw[hi][1]
                hi += 1
        return pl

    # ------------------------------------------------------------------ backward
    def run_backward(self, pl, dy, priors=None, mult=None):
        """dy: gradient w.r.t. plan['y'] ([B,P,5] fp32).  Fills self.gflat (overwrites)."""
        assert pl["train"]
        self.gflat.zero_()
        self.gpad.zero_()
        self.dwp.zero_()
        nb = len(self.blocks)
        drop_all = pl["drop"]
        offs, o = [], 0
        for (hh, ww) in self.head_hw:
            offs.append(o)
            o += hh * ww
        # ---- the four heads first: their input gradients only need dy and the saved block outputs, and the block loop
        # below adds head_dx of scale k-1 while it processes continue block k
        for hi in range(4):
            i = self.n_fe + hi
            w = self._view(self.pflat, f"extracting_layers.{hi}.0.weight")
            ops.ssd_head_bwd(pl["blocks"][i]["out"], pl["head_dx"][hi], w, mult, offs[hi], pl["y"], dy,
                             self._view(self.gflat, f"extracting_layers.{hi}.0.weight"),
                             self._view(self.gflat, f"extracting_layers.{hi}.0.bias"))
        # Weight-gradient launches go to side streams (parallel branches of the captured graph): they only read a block's
        # saved activations and its finished gp1 / gp2 / GS planes and accumulate into disjoint blocks of dwp / the bias
        # gradients, so nothing downstream waits for them until the final unpack -- at 16 images most of them are
        # latency-bound launches of a few CTAs (74 of them took 1.3 ms of a 3.4 ms step when serialised on one stream).
        main_stream = torch.cuda.current_stream()
        if getattr(self, "_wg_sides", None) is None:
            n_side = int(os.environ.get("FD_SSD_WGRAD_STREAMS", "4"))
            self._wg_sides = [torch.cuda.Stream(device=dy.device) for _ in range(n_side)]
        # only as branches of a captured graph: in eager mode the extra event records / waits cost more host time than the
        # overlap returns (the eager 2-GPU step, which is launch-bound, went from 4.1 to 5.3 ms with them)
        sides, rr = (self._wg_sides if torch.cuda.is_current_stream_capturing() else []), [0]

        def side():
            if not sides:
                return contextlib.nullcontext()       # eager: no stream switch at all (a context enter costs host time)
            rr[0] += 1
            return torch.cuda.stream(sides[rr[0] % len(sides)])

        for i in range(nb - 1, -1, -1):
            b, d = self.blocks[i], pl["blocks"][i]
            drop = drop_all[i] if drop_all is not None else [None] * b.go
            cur = d["inp"]
            # ---- gradient w.r.t. the block output: the last block only has its head; every other G already holds
            # (next block's input gradient [+ head dx of this scale]), written by the next block's step below
            G = pl["head_dx"][3] if i == nb - 1 else d["G"]
            # ---- through pool / dropout / LeakyReLU' of conv2
            if b.pool:
                for g in range(b.go):
                    ops.maxpool2x2_bwd(d["s"][g], G[g], d["gs"][g], d["mb"][g], drop[g], self.slope, d["gp2"][g],
                                       argmax=d["amax"][g])
                GS = d["gs"]
            else:
                for g in range(b.go):
                    ops.grad_mask(G[g], self.slope, d["mb"][g], drop[g], d["gp2"][g])
                GS = G
            # ---- gp1[h] = (sum_g dgrad(gp2[g], W2[g][h])) * lrelu'(a[h])
            pre2, pre1, pres = b.name + ".conv2", b.name + ".conv1", b.name + ".pointwise_conv_skip"
            wide2 = self.use_wide and pre2 in self.wide_d
            for hh in range(b.go // 2 if wide2 else 0):
                ops.conv3x3_wide(d["gp2"], self.wide_d[pre2][hh], slope=self.slope, mask_in=self._pair(d["ma"], hh),
                                 out2=self._pair(d["gp1"], hh))
            for h in range(0 if wide2 else b.go):
                prev = None
                for g in range(b.go):
                    wd = self.w_dgrad[self._sub(pre2, g, h)]
                    if g < b.go - 1:
                        dst = d["U"][h] if (g % 2 == 0) else d["U2"][h]
                        ops.conv3x3(d["gp2"][g], wd, slope=self.slope, residual=prev, out=dst)
                        prev = dst
                    else:
                        ops.conv3x3(d["gp2"][g], wd, slope=self.slope, residual=prev, mask_in=d["ma"][h], out2=d["gp1"][h])
            # ---- gradient w.r.t. the block input: conv1 dgrad + skip path (+ the previous scale's head gradient)
            if i > 0:
                pb, pd = self.blocks[i - 1], pl["blocks"][i - 1]
                target = pd["G"]
                extra = pl["head_dx"][i - 1 - self.n_fe] if (i - 1) >= self.n_fe else None
            else:
                target, extra = pl["g_stem"], None
            wide1 = self.use_wide and pre1 in self.wide_d and (not b.has_skip_conv or pres in self.wide_d)
            for hh in range(b.gi // 2 if wide1 else 0):
                final = self._pair(target, hh) if extra is None else self._pair(d["U"], hh)
                if b.has_skip_conv:
                    ops.conv3x3_wide(d["gp1"], self.wide_d[pre1][hh], slope=self.slope, out=self._pair(d["U2"], hh))
                    ops.conv3x3_wide(GS, self.wide_d[pres][hh], slope=self.slope, residual=self._pair(d["U2"], hh), out=final,
                                     flags=ops.CONV_1X1)
                else:
                    ops.conv3x3_wide(d["gp1"], self.wide_d[pre1][hh], slope=self.slope, residual=self._pair(GS, hh), out=final)
                if extra is not None:       # + head gradient of the previous scale
                    for h in (2 * hh, 2 * hh + 1):
                        ops.act_mask(d["U"][h], 1.0, None, extra[h], None, target[h])
            for h in range(0 if wide1 else b.gi):
                chain = [(d["gp1"][g], self.w_dgrad[self._sub(pre1, g, h)], 0) for g in range(b.go)]
                if b.has_skip_conv:
                    chain += [(GS[g], self.w_dgrad[self._sub(pres, g, h)], ops.CONV_1X1) for g in range(b.go)]
                prev = None
                if not b.has_skip_conv:
                    prev = GS[h]
                for k, (src, wd, fl) in enumerate(chain):
                    lastc = k == len(chain) - 1
                    if lastc and extra is None:
                        dst = target[h]
                    else:
                        dst = d["U"][h] if (k % 2 == 0) else d["U2"][h]
                    ops.conv3x3(src, wd, slope=self.slope, residual=prev, out=dst, flags=fl)
                    prev = dst
                if extra is not None:       # + head gradient of the previous scale: G_prev = prev + head_dx
                    ops.act_mask(prev, 1.0, None, extra[h], None, target[h])
            # ---- weight / bias gradients of the block
            n3 = 9 * 64 * 64
            dwp_flat = self.dwp.view(-1)
            for s_ in sides:
                s_.wait_stream(main_stream)       # this block's gp1 / gp2 / GS are final

            def wgrad_wide(pre, xs, gs, gi_, go_):
                """128 x 128 channel blocks of the weight gradient of layer `pre` (fd_conv3x3_wgrad_wide)"""
                for gg in range(go_ // 2):
                    for hh in range(gi_ // 2):
                        sub_off = [self._sub(pre, 2 * gg + c, 2 * hh + r) * n3 for r in range(2) for c in range(2)]
                        with side():
                            ops.conv3x3_wgrad_wide(xs[2 * hh], xs[2 * hh + 1], gs[2 * gg], gs[2 * gg + 1], dwp_flat, sub_off,
                                                   dbias0=self._gbias(pre, 2 * gg) if hh == 0 else None,
                                                   dbias1=self._gbias(pre, 2 * gg + 1) if hh == 0 else None)

            wide_w2 = self.use_wide and b.go % 2 == 0
            wide_w1 = wide_w2 and b.gi % 2 == 0
            if wide_w2:
                wgrad_wide(pre2, d["a"], d["gp2"], b.go, b.go)
            if wide_w1:
                wgrad_wide(pre1, cur, d["gp1"], b.gi, b.go)
                if b.has_skip_conv:
                    wgrad_wide(pres, cur, GS, b.gi, b.go)
            for g in range(b.go):
                for h in range(0 if wide_w2 else b.go):
                    with side():
                        ops.conv3x3_wgrad(d["a"][h], d["gp2"][g], self.dwp[self._sub(pre2, g, h)],
                                          self._gbias(pre2, g) if h == 0 else None)
                for h in range(0 if wide_w1 else b.gi):
                    with side():
                        ops.conv3x3_wgrad(cur[h], d["gp1"][g], self.dwp[self._sub(pre1, g, h)],
                                          self._gbias(pre1, g) if h == 0 else None)
                    if b.has_skip_conv:
                        with side():
                            ops.conv3x3_wgrad(cur[h], GS[g], self.dwp[self._sub(pres, g, h)],
                                              self._gbias(pres, g) if h == 0 else None)
        for s_ in sides:
            main_stream.wait_stream(s_)
        gsw = self.gpad[self.pad_stem_w_off:self.pad_stem_b_off].view(64, self.in_ch, 3, 3)
        gsb = self.gpad[self.pad_stem_b_off:self.pad_stem_b_off + 64]
        ops.stem_wgrad(pl["x"], pl["g_stem"][0], gsw, gsb, 2, 1)
        ops.unpack_wgrad3x3(self.dwp.view(self.n_sub, 9, 64, 64), self.gpad[:self.pad_b3_off].view(self.n_sub, 64, 64, 3, 3))
        # padded gradients -> compact conv list (gather) -> un-padded flat gradient buffer (scatter); heads are there already
        ops.index_copy(self.gconv, self.gpad, self.index_conv, scatter=False)
        ops.index_copy(self.gflat, self.gconv, self.conv_pos, scatter=True)

    # ------------------------------------------------------------------ fused train step (ModelMetaSSD.py:175)
    def train_step(self, x, gt, priors, mult, neg_pos_ratio=10, dropout=True, allreduce=None, optimizer=None,
                   num_pos_reduce=None):
        """forward -> ssd_loss(y_hat[:,:,0], y_hat[:,:,1:], y[:,:,0], y[:,:,1:], 10) with its gradient from the same
        kernel -> backward [-> all-reduce -> optimizer].  plan['loss'] = the loss (0-d), gflat = its gradient.
        ``num_pos_reduce``: sums the positive count over data-parallel ranks -- the gradient all-reduce then
        yields exactly the single-process gradient of the global batch."""
        with torch.cuda.device(x.device):
            pl = self.forward(x, train=True, dropout=dropout, priors=priors, mult=mult)
            y = pl["y"]
            conf = y[:, :, 0].contiguous()
            loc = y[:, :, 1:].contiguous()
            lab = gt[:, :, 0].contiguous()
            gl = gt[:, :, 1:].contiguous()
            ops.ssd_loss(conf, loc, lab, gl, int(neg_pos_ratio), pl["sums"], pl["npos"], None, pl["dconf"], pl["dloc"])
            n = pl["npos"].sum().float()
            if num_pos_reduce is not None:
                n = num_pos_reduce(n)
            torch.div(pl["sums"].sum(), n, out=pl["loss"])
            inv = 1.0 / n
            pl["dy"][:, :, 0] = pl["dconf"] * inv
            pl["dy"][:, :, 1:] = pl["dloc"] * inv
            self.run_backward(pl, pl["dy"], priors=priors, mult=mult)
            if allreduce is not None:
                allreduce(self.gflat)
            if optimizer is not None:
                optimizer.step()
        return pl
