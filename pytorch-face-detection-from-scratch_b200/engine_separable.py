"""Training engine of the depthwise-separable backbone (reference models/SeparableCNN.py:10-117): forward WITH saved
activations, backward and the train step, on G = filters / 64 NHWC bf16 channel planes.

The inference fast path of ``filters == 64`` stays the fused block kernel (``fd_sepblock_fwd``: both pointwise GEMMs and
the depthwise stage of a block in ONE launch, nothing through HBM).  Training needs the intermediates
(``t1 = lrelu(pw1 x)``, ``t2 = lrelu(dw t1)``) in memory anyway, so this engine runs the block layer by layer:

* pointwise convolutions: ``fd_conv3x3`` in centre-tap mode (``FD_CONV_1X1``) per (output plane, input plane) pair, the
  partial sums chained through the residual operand; forward and -- with the dgrad packing -- input gradient; weight
  gradients = the centre tap of ``fd_conv3x3_wgrad``;
* depthwise 3x3: ``fd_dwconv3x3_lrelu`` forward, ``fd_dwconv`` with flipped taps for the input gradient,
  ``fd_dwconv3x3_wgrad`` for the weight gradient;
* LeakyReLU': ``fd_lrelu_bwd`` (sign of the saved activation); Dropout2d multipliers, skip, pooling, stem and head as in
  the residual backbones (``fd_act_mask`` / ``fd_grad_mask`` / ``fd_maxpool2x2_*`` / ``fd_stem_*`` / ``fd_head_*``).

Parameters stay ordinary ``nn.Parameter`` tensors; gradients land in one flat fp32 buffer (``gflat``, the data-parallel
all-reduce unit) of which every ``p.grad`` is a view.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


class SeparableTrainEngine:
    def __init__(self, filters, in_ch, in_h, in_w, num_blocks, stem_k, stem_s, stem_pad, head_k, head_pad, block_patches,
                 slope=0.2, block_drop=0.25, head_drop=0.5):
        if filters % 64 != 0:
            raise NotImplementedError("SeparableCNN kernels exist for filters = 64 * G; got filters=%d" % filters)
        self.F, self.G, self.num_blocks, self.slope = filters, filters // 64, num_blocks, slope
        self.in_ch, self.in_h, self.in_w, self.stem_k = in_ch, in_h, in_w, stem_k
        self.stem_s, self.stem_pad, self.head_k, self.head_pad = stem_s, stem_pad, head_k, head_pad
        self.block_drop, self.head_drop = block_drop, head_drop
        H = (in_h + 2 * stem_pad - stem_k) // stem_s + 1
        W = (in_w + 2 * stem_pad - stem_k) // stem_s + 1
        self.H0, self.W0 = H, W
        self.shapes, self.pools = [], []
        for _ in range(num_blocks):
            self.shapes.append((H, W))
            pool = H > block_patches                       # SeparableCNN.py:49
            self.pools.append(pool)
            if pool:
                H, W = H // 2, W // 2
        self.So_h, self.So_w = H + 2 * head_pad - head_k + 1, W + 2 * head_pad - head_k + 1
        F_ = filters
        self.sections, off = {}, 0
        names = [("conv1.weight", (F_, in_ch, stem_k, stem_k)), ("conv1.bias", (F_,))]
        for k in range(num_blocks):
            pre = f"residual_blocks.{k}."
            names += [(pre + "pointwise_conv1.weight", (F_, F_, 1, 1)), (pre + "depthwise_conv.weight", (F_, 1, 3, 3)),
                      (pre + "pointwise_conv2.weight", (F_, F_, 1, 1))]
        names += [("out.weight", (5, F_, head_k, head_k)), ("out.bias", (5,))]
        for name, shape in names:
            n = 1
            for d in shape:
                n *= d
            self.sections[name] = (off, n, shape)
            off += (n + 3) // 4 * 4
        self.n_flat = off
        self.device, self.params, self.gflat = None, None, None
        self.plans: Dict[tuple, dict] = {}

    def param_names(self) -> List[str]:
        return list(self.sections.keys())

    def grad_view(self, name):
        off, n, shape = self.sections[name]
        return self.gflat[off:off + n].view(shape)

    def bind(self, params):
        dev = params["conv1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("the fd_b200 backbone runs on CUDA only (no CPU fallback): call model.cuda()")
        if dev != self.device or self.gflat is None:
            self.device = dev
            G, L = self.G, 2 * self.num_blocks
            nsub = L * G * G
            self.gflat = torch.zeros(self.n_flat, dtype=F32, device=dev)
            self.wpad = torch.zeros((nsub, 64, 64, 3, 3), dtype=F32, device=dev)      # only the centre tap is ever non-zero
            self.w_fwd = torch.empty((nsub, 9, 64, 64), dtype=BF16, device=dev)
            self.w_dgrad = torch.empty((nsub, 9, 64, 64), dtype=BF16, device=dev)
            self.dwp = torch.zeros((nsub, 9 * 64 * 64), dtype=F32, device=dev)
            self.dw_sub = torch.empty((nsub, 64, 64, 3, 3), dtype=F32, device=dev)
            self.zero64 = torch.zeros(64, dtype=F32, device=dev)
            self.plans.clear()
        self.params = params

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
        sub = torch.stack(pw).float().view(L, G, 64, G, 64).permute(0, 1, 3, 2, 4).reshape(L * G * G, 64, 64)
        self.wpad[:, :, :, 1, 1] = sub
        ops.pack_conv3x3(self.wpad, self.w_fwd, self.w_dgrad)
        d = torch.stack(dw).float().view(nb, G, 64, 9)
        self.w_dw = d.permute(0, 1, 3, 2).contiguous()                 # [nb, G, 9, 64] tap-major (fd_sep_pack layout)
        self.w_dw_flip = d.flip(-1).permute(0, 1, 3, 2).contiguous()   # taps reversed: the depthwise input gradient

    # ------------------------------------------------------------------ plan
    def plan(self, B):
        if B in self.plans:
            return self.plans[B]
        G, dev = self.G, self.device

        def planes(h, w):
            return [torch.empty((B, h, w, 64), dtype=BF16, device=dev) for _ in range(G)]

        pl = {"act0": planes(self.H0, self.W0), "g_stem": planes(self.H0, self.W0), "blocks": [], "generation": 0}
        pl["x_cache"] = ops.stem_cache(B, (self.in_ch, self.in_h, self.in_w), (self.stem_k, self.stem_s, self.stem_pad), dev)
        for (h, w), pool in zip(self.shapes, self.pools):
            b = {"T": planes(h, w), "T2": planes(h, w), "t1": planes(h, w), "t2": planes(h, w), "s": planes(h, w),
                 "gy": planes(h, w), "gu": planes(h, w), "gt1": planes(h, w), "gr1": planes(h, w)}
            ho, wo = (h // 2, w // 2) if pool else (h, w)
            b["out"] = planes(ho, wo) if pool else b["s"]
            b["G"] = planes(ho, wo)
            b["gs"] = planes(h, w) if pool else None
            b["amax"] = [torch.empty((B, ho, wo, 8), dtype=torch.int16, device=dev) for _ in range(G)] if pool else None
            pl["blocks"].append(b)
        pl["y"] = torch.empty((B, 5, self.So_h, self.So_w), dtype=F32, device=dev)
        pl["dy"] = torch.empty_like(pl["y"])
        pl["loss"] = torch.empty((B,), dtype=F32, device=dev)
        pl["drop"] = None
        self.plans[B] = pl
        return pl

    def _chain(self, srcs, wsel, layer, g, first_residual, dst_a, dst_b, final=None, transposed=False, **last_kw):
        """sum over the other plane index of centre-tap convolutions; forward: sum_h W[g][h] srcs[h]; transposed (input
        gradient): sum_h W[h][g]^T srcs[h] with the dgrad packing."""
        prev = first_residual
        dst = None
        for h in range(self.G):
            last = h == self.G - 1
            dst = final if (last and final is not None) else (dst_b if (h % 2) else dst_a)
            sub = self._sub(layer, h, g) if transposed else self._sub(layer, g, h)
            kw = dict(slope=self.slope, lrelu=False, residual=prev, out=dst, flags=ops.CONV_1X1)
            if last:
                kw.update(last_kw)
            ops.conv3x3(srcs[h], wsel[sub], **kw)
            prev = dst
        return dst

    # ------------------------------------------------------------------ forward (saves what the backward pass needs)
    def forward(self, x, dropout: bool):
        B = x.shape[0]
        G, nb, P = self.G, self.num_blocks, self.params
        pl = self.plan(B)
        pl["generation"] += 1
        self.pack_weights()
        if dropout:
            r = torch.rand((nb + 1, G, B, 64), device=x.device)
            scale = torch.empty_like(r)
            ops.dropout_scale(r, nb * G * B * 64, 1.0 - self.block_drop, 1.0 - self.head_drop, scale)
            pl["drop"] = scale
        else:
            pl["drop"] = None
        pl["x"] = x
        w1, b1 = P["conv1.weight"].detach().float(), P["conv1.bias"].detach().float()
        ops.stem_planes_fwd(x, w1, b1, pl["act0"], self.stem_s, self.stem_pad, x_cache=pl["x_cache"])
        cur = pl["act0"]
        for k, b in enumerate(pl["blocks"]):
            b["inp"] = cur
            drop = [pl["drop"][k, g] for g in range(G)] if pl["drop"] is not None else [None] * G
            for g in range(G):
                if G == 1:
                    ops.conv3x3(cur[0], self.w_fwd[self._sub(2 * k, 0, 0)], slope=self.slope, lrelu=True, out=b["t1"][0],
                                flags=ops.CONV_1X1)
                else:
                    raw = self._chain(cur, self.w_fwd, 2 * k, g, None, b["T"][g], b["T2"][g])
                    ops.act_mask(raw, self.slope, None, None, None, b["t1"][g])
                ops.dwconv3x3_lrelu(b["t1"][g], self.w_dw[k, g], self.slope, b["t2"][g])
            for g in range(G):
                if G == 1:
                    ops.conv3x3(b["t2"][0], self.w_fwd[self._sub(2 * k + 1, 0, 0)], slope=self.slope, lrelu=False,
                                chan_scale=drop[0], residual=cur[0], out=b["s"][0], flags=ops.CONV_1X1)
                else:
                    raw = self._chain(b["t2"], self.w_fwd, 2 * k + 1, g, None, b["T"][g], b["T2"][g])
                    ops.act_mask(raw, 1.0, drop[g], cur[g], None, b["s"][g])      # slope 1: no activation after pw2
                if self.pools[k]:
                    ops.maxpool2x2_fwd(b["s"][g], b["out"][g], b["amax"][g])
            cur = b["out"]
        pl["head_in"] = cur
        wo = P["out.weight"].detach().float()
        pl["w_head"] = [wo[:, g * 64:(g + 1) * 64].contiguous() for g in range(G)]
        hd = [pl["drop"][nb, g] for g in range(G)] if pl["drop"] is not None else [None] * G
        if G == 1:
            ops.head_fwd(cur[0], hd[0], pl["w_head"][0], P["out.bias"].detach().float(), pl["y"], self.head_pad)
        else:
            logits = None
            for g in range(G):
                wt = torch.empty(self.head_k * self.head_k * 5 * 64, dtype=F32, device=x.device)
                ops.head_pack(pl["w_head"][g], wt)
                part = torch.empty_like(pl["y"])
                ops.head_fwd(cur[g], hd[g], pl["w_head"][g], None, part, self.head_pad, w_t=wt)
                logits = part if logits is None else logits + part
            torch.sigmoid(logits + P["out.bias"].detach().float().view(1, 5, 1, 1), out=pl["y"])
        return pl

    # ------------------------------------------------------------------ backward
    def run_backward(self, pl, dy):
        G, nb, P = self.G, self.num_blocks, self.params
        drop_all = pl["drop"]
        self.gflat.zero_()
        self.dwp.zero_()
        last = pl["blocks"][nb - 1]
        gw_out = self.grad_view("out.weight")
        for g in range(G):
            dwg = torch.zeros_like(pl["w_head"][g])
            dbg = self.grad_view("out.bias") if g == 0 else torch.zeros(5, dtype=F32, device=dy.device)
            ops.head_bwd(pl["head_in"][g], drop_all[nb, g] if drop_all is not None else None, pl["w_head"][g], pl["y"], dy,
                         self.head_pad, last["G"][g], None, None, self.slope, None, dwg, dbg)
            gw_out[:, g * 64:(g + 1) * 64].copy_(dwg)
        for k in range(nb - 1, -1, -1):
            b = pl["blocks"][k]
            pre = f"residual_blocks.{k}."
            cur = b["inp"]
            drop = [drop_all[k, g] for g in range(G)] if drop_all is not None else [None] * G
            if self.pools[k]:
                for g in range(G):
                    ops.maxpool2x2_bwd(b["s"][g], b["G"][g], b["gs"][g], None, None, self.slope, None, argmax=b["amax"][g])
                GS = b["gs"]
            else:
                GS = b["G"]
            # gradient w.r.t. the pw2 output: through the Dropout2d multiplier (SeparableCNN.py:47-48)
            if drop_all is not None:
                for g in range(G):
                    ops.grad_mask(GS[g], self.slope, None, drop[g], b["gy"][g])
                GY = b["gy"]
            else:
                GY = GS
            gdw = self.grad_view(pre + "depthwise_conv.weight")
            for g in range(G):
                for h in range(G):
                    ops.conv3x3_wgrad(b["t2"][h], GY[g], self.dwp[self._sub(2 * k + 1, g, h)], None)
            for h in range(G):
                gt2 = self._chain(GY, self.w_dgrad, 2 * k + 1, h, None, b["T"][h], b["T2"][h], transposed=True)
                ops.lrelu_bwd(gt2, b["t2"][h], self.slope, b["gu"][h])
                ops.dwconv3x3_wgrad(b["t1"][h], b["gu"][h], gdw[h * 64:(h + 1) * 64])
                ops.dwconv(b["gu"][h], self.w_dw_flip[k, h], self.zero64, 3, 1, 1, 1, ops.ACT_NONE, b["gt1"][h])
                ops.lrelu_bwd(b["gt1"][h], b["t1"][h], self.slope, b["gr1"][h])
            for g in range(G):
                for h in range(G):
                    ops.conv3x3_wgrad(cur[h], b["gr1"][g], self.dwp[self._sub(2 * k, g, h)], None)
            target = pl["blocks"][k - 1]["G"] if k > 0 else pl["g_stem"]
            for h in range(G):
                self._chain(b["gr1"], self.w_dgrad, 2 * k, h, GS[h], b["T"][h], b["T2"][h], final=target[h], transposed=True)
        gw1, gb1 = self.grad_view("conv1.weight"), self.grad_view("conv1.bias")
        ops.stem_planes_wgrad(pl["x"], pl["g_stem"], gw1, gb1, self.stem_s, self.stem_pad, x_cache=pl["x_cache"])
        L = 2 * nb
        ops.unpack_wgrad3x3(self.dwp.view(L * G * G, 9, 64, 64), self.dw_sub)
        centre = self.dw_sub[:, :, :, 1, 1].view(L, G, G, 64, 64).permute(0, 1, 3, 2, 4).reshape(L, self.F, self.F)
        for k in range(nb):
            pre = f"residual_blocks.{k}."
            self.grad_view(pre + "pointwise_conv1.weight").copy_(centre[2 * k].view(self.F, self.F, 1, 1))
            self.grad_view(pre + "pointwise_conv2.weight").copy_(centre[2 * k + 1].view(self.F, self.F, 1, 1))

    def train_step(self, x, gt, dropout=True, allreduce=None, optimizer=None):
        with torch.cuda.device(x.device):
            pl = self.forward(x, dropout)
            if tuple(gt.shape) != tuple(pl["y"].shape):
                raise ValueError(f"target map {tuple(gt.shape)} does not match the head {tuple(pl['y'].shape)}")
            ops.yolo_loss(pl["y"], gt, pl["loss"], None, pl["dy"])
            self.run_backward(pl, pl["dy"])
            if allreduce is not None:
                allreduce(self.gflat)
            if optimizer is not None:
                optimizer.step()
        return pl
