"""MobilenetV3Backbone -- mirror of the reference's ``models/MobilenetV3Backbone.py:11-60``: timm
``tf_mobilenetv3_small_100`` minus its last five children + ``Conv2d(576 -> 5, 3x3, pad 1)`` + sigmoid, inference path
(BASELINE config 4).

timm is not a dependency: the module tree below re-creates the parameter / buffer names of the timm graph
(``feature_extractor.0.weight``, ``feature_extractor.1.running_mean``, ``feature_extractor.3.<stage>.<block>.conv_pw.weight``
...) so that the ``state_dict`` of the official TorchScript archive loads with ``strict=True``; the sub-modules only HOLD
the parameters.  The arithmetic runs in ``MobilenetV3Engine``: BatchNorm (eval, eps 1e-3) folded into the preceding
convolution, 1x1 convolutions as tcgen05 GEMMs (``fd_pw_conv``), depthwise 3x3 / 5x5 TF-"SAME" convolutions with the
SqueezeExcite channel sums fused (``fd_dwconv``), gate (``fd_se_gate``), ``fd_head3x3_fwd`` -- NHWC bf16 activations with
the network's own channel counts.  There is no CPU path and no training path (``pretrained=True`` of the reference
downloads ImageNet weights; here weights come from ``load_state_dict``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .BaseModel import BaseModel

BF16, F32 = torch.bfloat16, torch.float32
BN_EPS = 1e-3

# (kind, kernel, stride, expanded channels, out channels, activation, SE reduced channels or 0) per stage -- the
# tf_mobilenetv3_small_100 graph of the archive (code/__torch__/timm/models/efficientnet_blocks.py)
STAGES = (
    (("ds", 3, 2, 16, 16, "relu", 8),),
    (("ir", 3, 2, 72, 24, "relu", 0), ("ir", 3, 1, 88, 24, "relu", 0)),
    (("ir", 5, 2, 96, 40, "hswish", 24), ("ir", 5, 1, 240, 40, "hswish", 64), ("ir", 5, 1, 240, 40, "hswish", 64)),
    (("ir", 5, 1, 120, 48, "hswish", 32), ("ir", 5, 1, 144, 48, "hswish", 40)),
    (("ir", 5, 2, 288, 96, "hswish", 72), ("ir", 5, 1, 576, 96, "hswish", 144), ("ir", 5, 1, 576, 96, "hswish", 144)),
    (("cba", 1, 1, 576, 576, "hswish", 0),),
)
_ACT = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "hswish": ops.ACT_HSWISH}


def _bn(c):
    return nn.BatchNorm2d(c, eps=BN_EPS)


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - the engine runs the graph
        raise RuntimeError("executed by MobilenetV3Engine (CUDA only); call the parent model")


class SqueezeExcite(_Holder):
    def __init__(self, c, r):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, r, 1)
        self.conv_expand = nn.Conv2d(r, c, 1)


class DepthwiseSeparableConv(_Holder):
    def __init__(self, cin, cout, k, se):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, k, groups=cin, bias=False)
        self.bn1 = _bn(cin)
        if se:
            self.se = SqueezeExcite(cin, se)
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = _bn(cout)


class InvertedResidual(_Holder):
    def __init__(self, cin, cexp, cout, k, se):
        super().__init__()
        self.conv_pw = nn.Conv2d(cin, cexp, 1, bias=False)
        self.bn1 = _bn(cexp)
        self.conv_dw = nn.Conv2d(cexp, cexp, k, groups=cexp, bias=False)
        self.bn2 = _bn(cexp)
        if se:
            self.se = SqueezeExcite(cexp, se)
        self.conv_pwl = nn.Conv2d(cexp, cout, 1, bias=False)
        self.bn3 = _bn(cout)


class ConvBnAct(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn1 = _bn(cout)


def _same_pad_lo(size, k, s):
    """TF 'SAME' (timm layers/padding.py): total = max((ceil(i/s)-1)*s + k - i, 0); leading pad = total // 2."""
    total = max((-(-size // s) - 1) * s + k - size, 0)
    return total // 2


class MobilenetV3Engine:
    """Inference executor.  ``prepare(state)`` folds BatchNorm and packs every weight once per weight version;
    ``forward(x)`` issues the kernel sequence (fp32 / uint8 NCHW images -> sigmoid head ``[B,5,H/32,W/32]`` fp32)."""

    def __init__(self):
        self.device = None
        self.layers = None
        self.plans = {}
        self.version = None

    # ------------------------------------------------------------------ weights
    @staticmethod
    def _fold(sd, pre):
        g, b = sd[pre + "weight"].float(), sd[pre + "bias"].float()
        m, v = sd[pre + "running_mean"].float(), sd[pre + "running_var"].float()
        scale = g / torch.sqrt(v + BN_EPS)
        return scale.contiguous(), (b - m * scale).contiguous()

    def _pw(self, w, scale, bias):
        N, K = w.shape[0], w.shape[1]
        w2 = w.float().reshape(N, K).contiguous()
        packed = torch.empty(ops.pw_packed_elems(N, K), dtype=BF16, device=w.device)
        ops.pw_pack(w2, scale, packed)
        bpad = torch.zeros(ops.pw_padded_n(N), dtype=F32, device=w.device)
        bpad[:N] = bias
        return {"w": packed, "b": bpad, "N": N, "K": K}

    def _dw(self, w, scale, bias):
        C, K = w.shape[0], w.shape[2]
        packed = torch.empty((K * K, C), dtype=F32, device=w.device)
        ops.dw_pack(w.float().contiguous(), scale, packed)
        return {"w": packed, "b": bias, "K": K, "C": C}

    def _se(self, sd, pre):
        w1 = sd[pre + "conv_reduce.weight"].float()
        w2 = sd[pre + "conv_expand.weight"].float()
        return {"w1": w1.reshape(w1.shape[0], w1.shape[1]).contiguous(), "b1": sd[pre + "conv_reduce.bias"].float().contiguous(),
                "w2": w2.reshape(w2.shape[0], w2.shape[1]).contiguous(), "b2": sd[pre + "conv_expand.bias"].float().contiguous()}

    def prepare(self, sd, device):
        """sd: name -> tensor ON ``device`` (the module's parameters and buffers).  One-time weight preparation: the
        BatchNorm fold is plain fp32 tensor arithmetic on 17 k scale / shift values, the layouts are written by
        ``fd_pw_pack`` / ``fd_dw_pack``."""
        self.device = device
        fe = "feature_extractor."
        s0, b0 = self._fold(sd, fe + "1.")
        L = {"stem_w": (sd[fe + "0.weight"].float() * s0.view(-1, 1, 1, 1)).contiguous(), "stem_b": b0, "blocks": []}
        for si, stage in enumerate(STAGES):
            for bi, (kind, k, s, cexp, cout, act, se) in enumerate(stage):
                pre = f"{fe}3.{si}.{bi}."
                blk = {"kind": kind, "k": k, "s": s, "cexp": cexp, "cout": cout, "act": _ACT[act]}
                if kind == "cba":
                    sc, bb = self._fold(sd, pre + "bn1.")
                    blk["pw"] = self._pw(sd[pre + "conv.weight"], sc, bb)
                elif kind == "ds":
                    sc, bb = self._fold(sd, pre + "bn1.")
                    blk["dw"] = self._dw(sd[pre + "conv_dw.weight"], sc, bb)
                    sc, bb = self._fold(sd, pre + "bn2.")
                    blk["pwl"] = self._pw(sd[pre + "conv_pw.weight"], sc, bb)
                else:
                    sc, bb = self._fold(sd, pre + "bn1.")
                    blk["pw"] = self._pw(sd[pre + "conv_pw.weight"], sc, bb)
                    sc, bb = self._fold(sd, pre + "bn2.")
                    blk["dw"] = self._dw(sd[pre + "conv_dw.weight"], sc, bb)
                    sc, bb = self._fold(sd, pre + "bn3.")
                    blk["pwl"] = self._pw(sd[pre + "conv_pwl.weight"], sc, bb)
                if se:
                    blk["se"] = self._se(sd, pre + "se.")
                L["blocks"].append(blk)
        L["head_w"] = sd["out.weight"].float().contiguous()
        L["head_b"] = sd["out.bias"].float().contiguous()
        self.layers = L
        self.plans.clear()

    # ------------------------------------------------------------------ activation plan
    def _plan(self, B, H, W):
        key = (B, H, W)
        if key in self.plans:
            return self.plans[key]
        dev = self.device

        def buf(h, w, c):
            return torch.empty((B, h, w, c), dtype=BF16, device=dev)

        def out_size(size, k, s):
            return -(-size // s) if s > 1 else size

        pl = {"steps": []}
        h, w = -(-H // 2), -(-W // 2)
        pl["stem"] = buf(h, w, 16)
        pl["stem_pad"] = (_same_pad_lo(H, 3, 2), _same_pad_lo(W, 3, 2))
        cin = 16
        for blk in self.layers["blocks"]:
            st = {"blk": blk}
            k, s = blk["k"], blk["s"]
            if blk["kind"] == "cba":
                st["out"] = buf(h, w, blk["cout"])
            else:
                cdw = cin if blk["kind"] == "ds" else blk["cexp"]
                if blk["kind"] == "ir":
                    st["exp"] = buf(h, w, cdw)
                ho, wo = out_size(h, k, s), out_size(w, k, s)
                st["pad"] = (_same_pad_lo(h, k, s), _same_pad_lo(w, k, s)) if s > 1 else (k // 2, k // 2)
                st["dw"] = buf(ho, wo, cdw)
                if "se" in blk:
                    st["sum"] = torch.empty((B, ops.dwconv_se_blocks(ho, wo, cdw), cdw), dtype=F32, device=dev)
                    st["gate"] = torch.empty((B, cdw), dtype=F32, device=dev)
                st["out"] = buf(ho, wo, blk["cout"])
                st["skip"] = s == 1 and cin == blk["cout"]
                h, w = ho, wo
            cin = blk["cout"]
            pl["steps"].append(st)
        pl["y"] = torch.empty((B, 5, h, w), dtype=F32, device=dev)
        self.plans[key] = pl
        return pl

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, _, H, W = x.shape
        L = self.layers
        pl = self._plan(B, H, W)
        ops.mbv3_stem(x, L["stem_w"], L["stem_b"], pl["stem_pad"][0], pl["stem_pad"][1], pl["stem"])
        cur = pl["stem"]
        for st in pl["steps"]:
            blk = st["blk"]
            if blk["kind"] == "cba":
                ops.pw_conv(cur, blk["pw"]["w"], blk["pw"]["b"], blk["pw"]["N"], blk["act"], st["out"])
                cur = st["out"]
                continue
            t = cur
            if blk["kind"] == "ir":
                ops.pw_conv(cur, blk["pw"]["w"], blk["pw"]["b"], blk["pw"]["N"], blk["act"], st["exp"])
                t = st["exp"]
            ops.dwconv(t, blk["dw"]["w"], blk["dw"]["b"], blk["k"], blk["s"], st["pad"][0], st["pad"][1], blk["act"],
                       st["dw"], se_partial=st.get("sum"))
            if "se" in blk:
                se = blk["se"]
                ops.se_gate(st["sum"], st["dw"].shape[1] * st["dw"].shape[2], se["w1"], se["b1"], se["w2"], se["b2"],
                            st["gate"])
                ops.scale_channels(st["dw"], st["gate"])
            ops.pw_conv(st["dw"], blk["pwl"]["w"], blk["pwl"]["b"], blk["pwl"]["N"], ops.ACT_NONE, st["out"],
                        residual=cur if st["skip"] else None)
            cur = st["out"]
        ops.head3x3_fwd(cur, L["head_w"], L["head_b"], pl["y"])
        return pl["y"]


class MobilenetV3Backbone(BaseModel):
    def __init__(self, filters, input_shape, num_of_patches, probability_threshold=0.5, iou_threshold=0.5,
                 pretrained=True, input_kernel_size=10, input_stride=8, output_kernel_size=3, output_padding=0):
        super().__init__(filters, input_shape, num_of_patches=num_of_patches,
                         probability_threshold=probability_threshold, iou_threshold=iou_threshold)
        self.pretrained = pretrained          # no network here: weights come from load_state_dict (MobilenetV3Backbone.py:33-37)
        stages = []
        cin = 16
        for stage in STAGES:
            blocks = []
            for (kind, k, s, cexp, cout, act, se) in stage:
                if kind == "ds":
                    blocks.append(DepthwiseSeparableConv(cin, cout, k, se))
                elif kind == "ir":
                    blocks.append(InvertedResidual(cin, cexp, cout, k, se))
                else:
                    blocks.append(ConvBnAct(cin, cout))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        # children 0..3 of timm's model: conv_stem, bn1, act1, blocks (MobilenetV3Backbone.py:33-39 `[:-5]`)
        self.feature_extractor = nn.Sequential(nn.Conv2d(3, 16, 3, stride=2, bias=False), _bn(16), nn.Hardswish(),
                                               nn.Sequential(*stages))
        if output_kernel_size != 3:
            raise NotImplementedError("the fd_b200 MobilenetV3 head is the reference's 3x3 pad-1 convolution")
        self.out = nn.Conv2d(576, 5, stride=(1, 1), kernel_size=(output_kernel_size, output_kernel_size), padding=1)
        self.sigmoid = nn.Sigmoid()
        self.engine = MobilenetV3Engine()

    def _state_version(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def forward(self, x: torch.Tensor, predict: torch.Tensor = torch.tensor(0)):
        is_predict = bool(predict == 1)
        if is_predict:
            x = self._resize(self._to_model_device(x))            # MobilenetV3Backbone.py:51-54
            if x.dtype != torch.uint8:
                x = x / 255.0
            if len(x.shape) == 3:
                x = torch.unsqueeze(x, 0)
        if not x.is_cuda:
            raise RuntimeError("fd_b200 models run on CUDA tensors only (no CPU fallback)")
        if self.training:
            raise NotImplementedError("the fd_b200 MobilenetV3Backbone is the inference path (BatchNorm folded, eval "
                                      "statistics): call model.eval()")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        x = x.contiguous()
        with torch.no_grad(), torch.cuda.device(x.device):
            ver = self._state_version()
            if self.engine.version != ver or self.engine.device != x.device:
                sd = {k: v.detach() for k, v in self.state_dict().items()}
                self.engine.prepare(sd, x.device)
                self.engine.version = ver
            y = self.engine.forward(x).clone()
        if is_predict:
            return self.single_non_max_suppression(y[0])          # :59-60
        return y
