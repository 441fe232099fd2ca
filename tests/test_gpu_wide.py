"""GPU parity of fd_conv3x3_wide (64*gin -> 128 channels on [B,H,W,64] planes, tcgen05.mma.cta_group::2, weights streamed
from L2) through the C ABI against torch fp32 on the same bf16-rounded operands (models/PoolResnet.py:35-40 at
filters = 128, train_model.py:17; the 128 / 256-channel blocks of models/SSD.py:164-189).  Tolerances as in
test_gpu_layers.py: outputs are bf16 (rel. 2^-8 per element), accumulation fp32 over up to 9*256 products."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import fd, rel_err, require_cuda
from tests.test_gpu_layers import _bits, _close

pytestmark = pytest.mark.gpu


def _planes(t):
    """[B,H,W,64*G] -> list of G contiguous [B,H,W,64] planes"""
    return [t[..., 64 * g:64 * (g + 1)].contiguous() for g in range(t.shape[-1] // 64)]


def _cat(planes):
    return torch.cat([p.float() for p in planes], dim=-1)


@pytest.mark.parametrize("B,H,W,gin", [(2, 15, 15, 2), (1, 15, 15, 2), (3, 30, 30, 2), (2, 60, 60, 2), (1, 7, 9, 1), (5, 16, 8, 4),
                                       (1, 64, 125, 2), (150, 15, 15, 2), (9, 60, 60, 1)])
def test_conv3x3_wide_forward_and_dgrad(B, H, W, gin):
    require_cuda()
    ops = fd().ops
    torch.manual_seed(B * 100 + H + gin)
    dev = "cuda"
    Cin, Cout = 64 * gin, 128
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    res = torch.randn(B, H, W, Cout, device=dev).bfloat16()
    msk = torch.randn(B, H, W, Cout, device=dev).bfloat16()
    w = torch.randn(Cout, Cin, 3, 3, device=dev) * (0.05 / gin ** 0.5)
    bias = torch.randn(Cout, device=dev)
    cs = (torch.rand(B, Cout, device=dev) < 0.75).float() / 0.75
    cs2 = (torch.rand(B, Cout, device=dev) < 0.75).float() / 0.75
    wf = torch.empty(1, 1, gin, 9, 128, 64, dtype=torch.bfloat16, device=dev)
    ops.pack_conv3x3_wide(w, wf, None)
    wq = w.bfloat16().float()
    xp = _planes(x)
    xn = x.float().permute(0, 3, 1, 2)
    nhwc = lambda t: t.permute(0, 2, 3, 1)
    csp = [cs[:, :64].contiguous(), cs[:, 64:].contiguous()]
    cs2p = [cs2[:, :64].contiguous(), cs2[:, 64:].contiguous()]
    # (1) conv + bias + lrelu
    out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ops.conv3x3_wide(xp, wf[0, 0], bias=bias, lrelu=True, out=out)
    ref1 = nhwc(F.leaky_relu(F.conv2d(xn, wq, bias, padding=1), 0.2))
    _close(_cat(out), ref1, "conv+bias+lrelu")
    # (2) + dropout multiplier, sign bits of the pre-residual value, residual add
    mo = [torch.zeros(B, H, W, 2, dtype=torch.int32, device=dev) for _ in range(2)]
    out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ops.conv3x3_wide(xp, wf[0, 0], bias=bias, lrelu=True, chan_scale=csp, residual=_planes(res), mask_out=mo, out=out)
    refb = ref1 * cs[:, None, None, :]
    _close(_cat(out), refb + res.float(), "residual out")
    got_bits = torch.cat([((m.to(torch.int64)[..., None] >> torch.arange(32, device=dev)) & 1).reshape(B, H, W, 64) for m in mo],
                         dim=-1).bool()
    sure = refb.abs() > 1e-3
    assert torch.equal(got_bits[sure], (refb > 0)[sure]) and sure.float().mean().item() > 0.5
    # (3) the input gradient: a convolution Cout -> Cin with the dgrad packing (needs Cin = 128 output channels here)
    if gin == 2:
        g = torch.randn(B, H, W, Cout, device=dev).bfloat16()
        wd = torch.empty(1, 1, Cout // 64, 9, 128, 64, dtype=torch.bfloat16, device=dev)
        ops.pack_conv3x3_wide(w, None, wd)
        xin = torch.zeros(B, Cin, H, W, device=dev, requires_grad=True)
        (gref,) = torch.autograd.grad(F.conv2d(xin, wq, None, padding=1), xin, g.float().permute(0, 3, 1, 2))
        G = nhwc(gref) + res.float()
        out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        ops.conv3x3_wide(_planes(g), wd[0, 0], residual=_planes(res), out=out)
        _close(_cat(out), G, "dgrad + residual")
        out2 = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        ops.conv3x3_wide(_planes(g), wd[0, 0], residual=_planes(res), mask_in=[_bits(m) for m in _planes(msk)], chan_scale2=cs2p,
                         out2=out2)
        G2 = G * cs2[:, None, None, :] * torch.where(msk.float() > 0, 1.0, 0.2)
        _close(_cat(out2), G2, "masked out2", rel=6e-3)
        # shared-tile mode (fewer two-block tiles than SM pairs) writes BOTH outputs in one launch, bit-identical
        if ops.conv3x3_wide_shared_tile(B, H, W):
            oa = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
            ob = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
            ops.conv3x3_wide(_planes(g), wd[0, 0], residual=_planes(res), out=oa, mask_in=[_bits(m) for m in _planes(msk)],
                             chan_scale2=cs2p, out2=ob)
            assert all(torch.equal(a, b) for a, b in zip(oa, out)) and all(torch.equal(a, b) for a, b in zip(ob, out2))


def test_conv3x3_wide_centre_tap_is_pointwise():
    """FD_CONV_1X1: only the centre tap of the packed weights = a 1x1 convolution (models/SeparableCNN.py:13-19 at 128 filters)"""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(5)
    dev = "cuda"
    B, H, W = 4, 30, 30
    x = torch.randn(B, H, W, 128, device=dev).bfloat16()
    w1 = torch.randn(128, 128, device=dev) * 0.1
    w = torch.zeros(128, 128, 3, 3, device=dev)
    w[:, :, 1, 1] = w1
    wf = torch.empty(1, 1, 2, 9, 128, 64, dtype=torch.bfloat16, device=dev)
    ops.pack_conv3x3_wide(w, wf, None)
    out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ops.conv3x3_wide(_planes(x), wf[0, 0], lrelu=True, out=out, flags=ops.CONV_1X1)
    ref = F.leaky_relu(x.float() @ w1.bfloat16().float().t(), 0.2)
    _close(_cat(out), ref, "1x1")


def test_conv3x3_wide_256_outputs_two_groups():
    """Cout = 256 = two launches over the two 128-cout groups of the packing (models/SSD.py:173-181: 128 -> 256)"""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(6)
    dev = "cuda"
    B, H, W = 2, 30, 30
    x = torch.randn(B, H, W, 128, device=dev).bfloat16()
    w = torch.randn(256, 128, 3, 3, device=dev) * 0.04
    wf = torch.empty(1, 2, 2, 9, 128, 64, dtype=torch.bfloat16, device=dev)
    ops.pack_conv3x3_wide(w, wf, None)
    outs = []
    for go in range(2):
        out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        ops.conv3x3_wide(_planes(x), wf[0, go], out=out)
        outs += out
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), None, padding=1).permute(0, 2, 3, 1)
    _close(_cat(outs), ref, "256 outputs")


@pytest.mark.parametrize("nprob,B,H,W", [(1, 2, 15, 15), (3, 5, 15, 15), (1, 3, 30, 30), (2, 2, 60, 60), (1, 1, 64, 125), (1, 70, 15, 15)])
def test_conv3x3_wgrad_wide_vs_autograd(nprob, B, H, W):
    """fd_conv3x3_wgrad_wide (one 128 x 128 channel block of dW per call, tcgen05.mma.cta_group::2, taps 0..7 + tap 8 in two
    passes, TMA reduce-stores) against the weight / bias gradient of torch conv2d in fp32 on the bf16-rounded operands;
    accumulates on top of what the buffers hold."""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(nprob * 1000 + B * 10 + H)
    dev = "cuda"
    x = torch.randn(nprob, B, H, W, 128, device=dev).bfloat16()
    g = (torch.randn(nprob, B, H, W, 128, device=dev) * 0.5).bfloat16()
    xp = [x[..., :64].contiguous(), x[..., 64:].contiguous()]
    gp = [g[..., :64].contiguous(), g[..., 64:].contiguous()]
    n3 = 9 * 64 * 64
    # packed sub-blocks per problem: [(g plane c, x plane r)] -> index c * 2 + r (the PlanarEngine order (g, h))
    dwp = torch.full((nprob, 4, n3), 0.5, dtype=torch.float32, device=dev)
    db = torch.full((nprob, 2, 64), 0.25, dtype=torch.float32, device=dev)
    sub_off = [(c * 2 + r) * n3 for r in range(2) for c in range(2)]
    ops.conv3x3_wgrad_wide(xp[0], xp[1], gp[0], gp[1], dwp.view(-1), sub_off, dw_stride=4 * n3, dbias0=db.view(-1),
                           dbias1=db.view(-1)[64:], dbias_stride=128)
    dw = torch.empty(nprob * 4, 64, 64, 3, 3, device=dev)
    ops.unpack_wgrad3x3((dwp - 0.5).view(nprob * 4, 9, 64, 64), dw)
    for q in range(nprob):
        w = torch.zeros(128, 128, 3, 3, device=dev, requires_grad=True)
        bias = torch.zeros(128, device=dev, requires_grad=True)
        y = F.conv2d(x[q].float().permute(0, 3, 1, 2), w, bias, padding=1)
        gw, gb = torch.autograd.grad(y, (w, bias), g[q].float().permute(0, 3, 1, 2))
        got = dw[q * 4:(q + 1) * 4].view(2, 2, 64, 64, 3, 3).permute(0, 2, 1, 3, 4, 5).reshape(128, 128, 3, 3)   # [(c,co),(r,ci)]
        assert rel_err(got, gw) <= 2e-3, (q, rel_err(got, gw))
        assert rel_err((db[q] - 0.25).reshape(-1), gb) <= 2e-3


@pytest.mark.parametrize("B,H,W,nblocks", [(3, 15, 15, 2), (64, 15, 15, 8), (5, 12, 14, 1)])
def test_conv3x3_wide_chain_is_bit_identical_to_single_launches(B, H, W, nblocks):
    """fd_conv3x3_wide_chain (a whole run of residual blocks in ONE launch, the pair keeps its image and re-loads what it just
    wrote) against the same layers as individual fd_conv3x3_wide launches: forward (models/PoolResnet.py:35-42 -- conv, bias,
    LeakyReLU, sign masks, Dropout2d multiplier, skip) and the input-gradient run (masked / scaled second outputs)."""
    require_cuda()
    ops = fd().ops
    if not ops.conv3x3_wide_chain_ok(B, H, W):
        pytest.skip("map does not fit a CTA pair")
    torch.manual_seed(B + H + nblocks)
    dev = "cuda"
    nl = 2 * nblocks
    bf = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)
    w = torch.randn(nl, 128, 128, 3, 3, device=dev) * 0.03
    bias = torch.randn(nl, 128, device=dev) * 0.1
    wf = torch.empty(nl, 1, 2, 9, 128, 64, dtype=torch.bfloat16, device=dev)
    wd = torch.empty(nl, 1, 2, 9, 128, 64, dtype=torch.bfloat16, device=dev)
    ops.pack_conv3x3_wide(w, wf, wd)
    drop = (torch.rand(nblocks, 2, B, 64, device=dev) < 0.75).float() / 0.75
    x0 = torch.randn(B, H, W, 128, device=dev).bfloat16()

    def forward(chain):
        XA = [bf(nl + 1, B, H, W, 64) for _ in range(2)]
        masks = [[torch.zeros(B, H, W, 2, dtype=torch.int32, device=dev) for _ in range(2)] for _ in range(nl)]
        for g in range(2):
            XA[g][0].copy_(x0[..., 64 * g:64 * g + 64])
        layers = []
        for i in range(nblocks):
            a = dict(bias=bias[2 * i], lrelu=True, mask_out=masks[2 * i], out=[XA[g][2 * i + 1] for g in range(2)])
            b = dict(bias=bias[2 * i + 1], lrelu=True, chan_scale=[drop[i, 0], drop[i, 1]], residual=[XA[g][2 * i] for g in range(2)],
                     mask_out=masks[2 * i + 1], out=[XA[g][2 * i + 2] for g in range(2)])
            if chain:
                layers += [ops.wide_chain_layer(2 * i, 2 * i, **a), ops.wide_chain_layer(2 * i + 1, 2 * i + 1, **b)]
            else:
                ops.conv3x3_wide([XA[g][2 * i] for g in range(2)], wf[2 * i, 0], **a)
                ops.conv3x3_wide([XA[g][2 * i + 1] for g in range(2)], wf[2 * i + 1, 0], **b)
        if chain:
            ops.conv3x3_wide_chain(XA, wf, layers)
        return XA, masks

    XA_c, m_c = forward(True)
    XA_s, m_s = forward(False)
    for g in range(2):
        assert torch.equal(XA_c[g].view(torch.int16), XA_s[g].view(torch.int16))
    # the layer hand-over (plain stores -> fence -> cluster barrier -> TMA loads of either CTA) must hold every time:
    # a visibility race would show up as an occasional stale halo row
    for rep in range(12 if B >= 64 else 3):
        XA_r, _ = forward(True)
        for g in range(2):
            assert torch.equal(XA_r[g].view(torch.int16), XA_s[g].view(torch.int16)), f"repetition {rep}"
    assert all(torch.equal(a, b) for la, lb in zip(m_c, m_s) for a, b in zip(la, lb))
    assert XA_s[0][nl].float().abs().max() > 0

    g_top = (torch.randn(B, H, W, 128, device=dev) * 0.1).bfloat16()

    def backward(chain):
        GP = [bf(nl, B, H, W, 64) for _ in range(2)]                    # [2 i] = gp1 of block i, [2 i + 1] = gp2
        Gb = [[bf(B, H, W, 64) for _ in range(2)] for _ in range(nblocks + 1)]      # [i + 1] = gradient of block i's output
        for g in range(2):
            Gb[nblocks][g].copy_(g_top[..., 64 * g:64 * g + 64])
            GP[g][nl - 1].copy_(g_top[..., 64 * g:64 * g + 64])         # stand-in for the head's masked gradient
        layers = []
        for i in range(nblocks - 1, -1, -1):
            d2 = dict(mask_in=m_s[2 * i], out2=[GP[g][2 * i] for g in range(2)])
            d1 = dict(residual=Gb[i + 1], out=Gb[i])
            if i > 0:
                d1.update(mask_in=m_s[2 * i - 1], chan_scale2=[drop[i - 1, 0], drop[i - 1, 1]], out2=[GP[g][2 * i - 1] for g in range(2)])
            if chain:
                layers += [ops.wide_chain_layer(2 * i + 1, 2 * i + 1, **d2), ops.wide_chain_layer(2 * i, 2 * i, **d1)]
            else:
                ops.conv3x3_wide([GP[g][2 * i + 1] for g in range(2)], wd[2 * i + 1, 0], **d2)
                if "out2" in d1 and not ops.conv3x3_wide_shared_tile(B, H, W):     # both outputs: shared-tile launches only
                    d1b = {k: v for k, v in d1.items() if k != "out"}
                    ops.conv3x3_wide([GP[g][2 * i] for g in range(2)], wd[2 * i, 0], **d1b)
                    d1 = dict(residual=d1["residual"], out=d1["out"])
                ops.conv3x3_wide([GP[g][2 * i] for g in range(2)], wd[2 * i, 0], **d1)
        if chain:
            ops.conv3x3_wide_chain(GP, wd, layers)
        return GP, Gb

    GP_c, G_c = backward(True)
    GP_s, G_s = backward(False)
    for g in range(2):
        assert torch.equal(GP_c[g].view(torch.int16), GP_s[g].view(torch.int16))
    assert all(torch.equal(a.view(torch.int16), b.view(torch.int16)) for la, lb in zip(G_c, G_s) for a, b in zip(la, lb))
    assert G_s[0][0].float().abs().max() > 0
