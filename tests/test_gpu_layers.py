"""GPU parity of the backbone kernels through the C ABI against torch fp32 on the same bf16-rounded
inputs.  Tolerances: outputs are stored as bf16 (rel. 2^-8 per element), accumulation is fp32, so
max-abs error <= 1e-2 * max|ref| and relative L2 error <= 4e-3 are required."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import fd, rel_err, require_cuda

pytestmark = pytest.mark.gpu
C = 64


def _bits(t):
    """sign-bit mask of a [B,H,W,C] tensor in the library's layout: int32 [B,H,W,C/32], bit c%32 of word c/32."""
    b = (t.float() > 0).to(torch.int64).reshape(*t.shape[:-1], t.shape[-1] // 32, 32)
    w = (b << torch.arange(32, device=t.device, dtype=torch.int64)).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()


def _close(got, ref, what, rel=4e-3, mx=1e-2):
    r = rel_err(got.float(), ref)
    m = (got.float() - ref).abs().max().item()
    assert r <= rel, f"{what}: rel L2 {r}"
    assert m <= mx * max(1.0, ref.abs().max().item()), f"{what}: max abs {m}"


@pytest.mark.parametrize("B,H,W", [(2, 15, 15), (3, 30, 30), (2, 60, 60), (1, 7, 9), (5, 16, 8), (2, 120, 120),
                                   (1, 64, 125), (160, 15, 15)])
def test_conv3x3_forward_epilogues(B, H, W):
    require_cuda()
    ops = fd().ops
    torch.manual_seed(B * 100 + H)
    dev = "cuda"
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    res = torch.randn(B, H, W, C, device=dev).bfloat16()
    msk = torch.randn(B, H, W, C, device=dev).bfloat16()
    w = torch.randn(C, C, 3, 3, device=dev) * 0.05
    bias = torch.randn(C, device=dev)
    cs = (torch.rand(B, C, device=dev) < 0.75).float() / 0.75
    cs2 = (torch.rand(B, C, device=dev) < 0.75).float() / 0.75
    wf = torch.empty(9, C, C, dtype=torch.bfloat16, device=dev); wd = torch.empty_like(wf)
    ops.pack_conv3x3(w, wf, wd)
    wq = w.bfloat16().float()
    xn = x.float().permute(0, 3, 1, 2)
    nhwc = lambda t: t.permute(0, 2, 3, 1)
    # (1) conv + bias + lrelu
    out = torch.zeros_like(x)
    ops.conv3x3(x, wf, bias=bias, lrelu=True, out=out)
    ref1 = nhwc(F.leaky_relu(F.conv2d(xn, wq, bias, padding=1), 0.2))
    _close(out, ref1, "conv+bias+lrelu")
    # (2) conv2 of the block: + dropout multiplier, sign bits of the pre-residual value, residual add
    mo = torch.zeros(B, H, W, C // 32, dtype=torch.int32, device=dev); out = torch.zeros_like(x)
    ops.conv3x3(x, wf, bias=bias, lrelu=True, chan_scale=cs, residual=res, mask_out=mo, out=out)
    refb = ref1 * cs[:, None, None, :]
    _close(out, refb + res.float(), "residual out")
    # sign bits, compared where the value is not within rounding distance of zero
    got_bits = ((mo.to(torch.int64)[..., None] >> torch.arange(32, device=dev)) & 1).reshape(B, H, W, C).bool()
    sure = refb.abs() > 1e-3
    assert torch.equal(got_bits[sure], (refb > 0)[sure]) and sure.float().mean().item() > 0.5
    # (3) dgrad packing + masked second output (backward chain): both outputs, then out2 alone (staged path)
    out = torch.zeros_like(x); out2 = torch.zeros_like(x)
    ops.conv3x3(x, wd, residual=res, out=out, mask_in=_bits(msk), chan_scale2=cs2, out2=out2)
    only2 = torch.zeros_like(x)
    ops.conv3x3(x, wd, residual=res, mask_in=_bits(msk), chan_scale2=cs2, out2=only2)
    assert torch.equal(only2, out2)
    xin = torch.zeros(B, C, H, W, device=dev, requires_grad=True)
    (gref,) = torch.autograd.grad(F.conv2d(xin, wq, None, padding=1), xin, xn)
    G = nhwc(gref) + res.float()
    _close(out, G, "dgrad + residual")
    G2 = G * cs2[:, None, None, :] * torch.where(msk.float() > 0, 1.0, 0.2)
    _close(out2, G2, "masked out2", rel=6e-3)


@pytest.mark.parametrize("B,H,W", [(3, 60, 60), (2, 30, 30), (5, 16, 8), (1, 240, 240), (2, 120, 120), (70, 30, 30), (1, 64, 126)])
def test_conv3x3_fused_pool_equals_conv_then_pool(B, H, W):
    """fd_conv3x3_pool (MaxPool2d(2) in the conv2 epilogue, models/PoolResnet.py:37-42) == fd_conv3x3 followed by
    fd_maxpool2x2_fwd, bit for bit: pooled activations, window positions and the sign-bit mask of the un-pooled value."""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(B * 31 + W)
    dev = "cuda"
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    res = torch.randn(B, H, W, C, device=dev).bfloat16()
    w = torch.randn(C, C, 3, 3, device=dev) * 0.05
    bias = torch.randn(C, device=dev)
    cs = (torch.rand(B, C, device=dev) < 0.75).float() / 0.75
    wf = torch.empty(9, C, C, dtype=torch.bfloat16, device=dev)
    ops.pack_conv3x3(w, wf, None)
    s_ = torch.zeros_like(x); m1 = torch.zeros(B, H, W, C // 32, dtype=torch.int32, device=dev)
    ops.conv3x3(x, wf, bias=bias, lrelu=True, chan_scale=cs, residual=res, mask_out=m1, out=s_)
    y1 = torch.zeros(B, H // 2, W // 2, C, dtype=torch.bfloat16, device=dev)
    a1 = torch.zeros(B, H // 2, W // 2, C // 8, dtype=torch.int16, device=dev)
    ops.maxpool2x2_fwd(s_, y1, a1)
    y2 = torch.full_like(y1, 3.0); a2 = torch.full_like(a1, -1); m2 = torch.zeros_like(m1)
    ops.conv3x3_pool(x, wf, y2, bias=bias, chan_scale=cs, residual=res, mask_out=m2, argmax=a2)
    assert torch.equal(y1, y2) and torch.equal(a1, a2) and torch.equal(m1, m2)
    y3 = torch.full_like(y1, 3.0)
    ops.conv3x3_pool(x, wf, y3, bias=bias, chan_scale=cs, residual=res)          # inference: no mask, no positions
    assert torch.equal(y1, y3)
    ref = F.max_pool2d(s_.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(y2.float(), ref)


@pytest.mark.parametrize("B,H,W", [(2, 15, 15), (3, 30, 30), (2, 60, 60), (1, 7, 9), (64, 15, 15), (2, 120, 120),
                                   (1, 64, 125), (1, 240, 240)])
def test_conv3x3_wgrad(B, H, W):
    require_cuda()
    ops = fd().ops
    torch.manual_seed(B + H)
    dev = "cuda"
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    g = (torch.randn(B, H, W, C, device=dev) * 0.1).bfloat16()
    dwp = torch.zeros(9, C, C, device=dev); db = torch.zeros(C, device=dev)
    ops.conv3x3_wgrad(x, g, dwp, db)
    dw = torch.empty(C, C, 3, 3, device=dev)
    ops.unpack_wgrad3x3(dwp, dw)
    wz = torch.zeros(C, C, 3, 3, device=dev, requires_grad=True)
    (dref,) = torch.autograd.grad(F.conv2d(x.float().permute(0, 3, 1, 2), wz, None, padding=1), wz,
                                  g.float().permute(0, 3, 1, 2))
    assert rel_err(dw, dref) <= 1e-4            # fp32 accumulate of exact bf16 products
    assert rel_err(db, g.float().sum(dim=(0, 1, 2))) <= 1e-4
    # accumulate semantics: a second call doubles
    ops.conv3x3_wgrad(x, g, dwp, db)
    ops.unpack_wgrad3x3(dwp, dw)
    assert rel_err(dw, 2 * dref) <= 1e-4


@pytest.mark.parametrize("u8", [False, True])
def test_stem_fwd_wgrad(u8):
    require_cuda()
    ops = fd().ops
    torch.manual_seed(5)
    dev = "cuda"
    B = 3
    if u8:
        x = torch.randint(0, 256, (B, 3, 480, 480), dtype=torch.uint8, device=dev)
        xf = x.float() / 255.0
    else:
        x = torch.rand(B, 3, 480, 480, device=dev)
        xf = x
    w = (torch.randn(C, 3, 10, 10, device=dev) * 0.05).requires_grad_(True)
    b = torch.randn(C, device=dev).requires_grad_(True)
    y = torch.zeros(B, 60, 60, C, dtype=torch.bfloat16, device=dev)
    ops.stem_fwd(x, w.detach(), b.detach(), y, 8, 2)
    ref = F.conv2d(xf, w, b, stride=8, padding=2)
    _close(y, ref.detach().permute(0, 2, 3, 1), "stem fwd")
    g = (torch.randn(B, 60, 60, C, device=dev) * 0.1).bfloat16()
    dw = torch.zeros(C, 3, 10, 10, device=dev); dbias = torch.zeros(C, device=dev)
    ops.stem_wgrad(x, g, dw, dbias, 8, 2)
    ref.backward(g.float().permute(0, 3, 1, 2))
    # the tensor-core stem rounds the image to bf16 on the way into shared memory: bf16 tolerance
    # against the fp32 reference, fp32-accumulation tolerance against the same conv on the rounded image
    assert rel_err(dw, w.grad) <= 4e-3 and rel_err(dbias, b.grad) <= 1e-4
    wz = torch.zeros_like(w).requires_grad_(True)
    (dref,) = torch.autograd.grad(F.conv2d(xf.bfloat16().float(), wz, None, stride=8, padding=2), wz,
                                  g.float().permute(0, 3, 1, 2))
    assert rel_err(dw, dref) <= 1e-4
    # bf16 image cache: the forward fills it, the weight gradient reads it instead of the fp32 / uint8 images
    n_cache = ops.stem_cache_elems(B, 3, 480, 480, C, 10, 8, 2)
    assert n_cache == 2 * 3 * 480 * 2 * 512          # B = 3 -> two image pairs, the odd slot stays zero
    cache = torch.zeros(n_cache, dtype=torch.bfloat16, device=dev)
    y2 = torch.zeros_like(y)
    ops.stem_fwd(x, w.detach(), b.detach(), y2, 8, 2, x_cache=cache)
    assert torch.equal(y2, y)
    want = torch.zeros(2, 3, 480, 2, 512, device=dev)
    xp = torch.cat([xf, torch.zeros(1, 3, 480, 480, device=dev)]).reshape(2, 2, 3, 480, 480)
    want[:, :, :, :, 2:482] = xp.permute(0, 2, 3, 1, 4)
    assert torch.equal(cache.view(2, 3, 480, 2, 512).float(), want.bfloat16().float())
    dw2 = torch.zeros_like(dw); dbias2 = torch.zeros_like(dbias)
    ops.stem_wgrad(x, g, dw2, dbias2, 8, 2, x_cache=cache)
    assert rel_err(dw2, dw) <= 1e-5 and rel_err(dbias2, dbias) <= 1e-5     # same MMAs, fp32 atomics reorder sums
    # second 64-channel plane of a 128-filter stem: forward from the bf16 copy, both weight gradients in one pass
    wb = (torch.randn(C, 3, 10, 10, device=dev) * 0.05)
    bb = torch.randn(C, device=dev)
    yb = torch.zeros_like(y); yb_ref = torch.zeros_like(y)
    ops.stem_fwd(x, wb, bb, yb_ref, 8, 2)
    ops.stem_fwd_cached(cache, x.shape, wb, bb, yb, 8, 2)
    assert torch.equal(yb, yb_ref)                                          # same A tile, same MMAs, same epilogue
    gb = (torch.randn(B, 60, 60, C, device=dev) * 0.1).bfloat16()
    dwb = torch.zeros_like(dw); dbiasb = torch.zeros_like(dbias)
    ops.stem_wgrad(x, gb, dwb, dbiasb, 8, 2, x_cache=cache)
    dwp = torch.zeros(2 * C, 3, 10, 10, device=dev); dbp = torch.zeros(2 * C, device=dev)
    ops.stem_wgrad_pair(cache, x.shape, g, gb, dwp, dbp, 8, 2)
    assert rel_err(dwp[:C], dw2) <= 1e-5 and rel_err(dwp[C:], dwb) <= 1e-5
    assert rel_err(dbp[:C], dbias2) <= 1e-5 and rel_err(dbp[C:], dbiasb) <= 1e-5
    ya = [torch.zeros_like(y), torch.zeros_like(y)]
    cache2 = torch.zeros_like(cache)
    ops.stem_planes_fwd(x, torch.cat([w.detach(), wb]), torch.cat([b.detach(), bb]), ya, 8, 2, x_cache=cache2)
    assert torch.equal(ya[0], y) and torch.equal(ya[1], yb_ref) and torch.equal(cache2, cache)
    dwq = torch.zeros_like(dwp); dbq = torch.zeros_like(dbp)
    ops.stem_planes_wgrad(x, [g, gb], dwq, dbq, 8, 2, x_cache=cache2)
    assert rel_err(dwq, dwp) <= 1e-5 and rel_err(dbq, dbp) <= 1e-5


def test_conv3x3_wgrad_multi_problem():
    """Eight stacked problems in one launch == eight single launches (strided outputs)."""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(3)
    dev = "cuda"
    P, B, H = 8, 4, 15
    x = torch.randn(P, B, H, H, C, device=dev).bfloat16()
    g = (torch.randn(P, B, H, H, C, device=dev) * 0.1).bfloat16()
    n3 = 9 * C * C
    dwp = torch.zeros(2 * P * n3, device=dev); db = torch.zeros(2 * P * C, device=dev)
    ops.conv3x3_wgrad_multi(x, g, dwp[n3:], 2 * n3, db[C:], 2 * C)      # odd layers, like conv2 of a run
    for q in range(P):
        one = torch.zeros(n3, device=dev); ob = torch.zeros(C, device=dev)
        ops.conv3x3_wgrad(x[q], g[q], one, ob)
        assert rel_err(dwp[(2 * q + 1) * n3:(2 * q + 2) * n3], one) <= 1e-5
        assert rel_err(db[(2 * q + 1) * C:(2 * q + 2) * C], ob) <= 1e-5
        assert dwp[(2 * q) * n3:(2 * q + 1) * n3].abs().max().item() == 0.0     # even slots untouched


@pytest.mark.parametrize("H,K,pad", [(15, 6, 0), (15, 3, 1)])
def test_head_fwd_bwd(H, K, pad):
    require_cuda()
    ops = fd().ops
    torch.manual_seed(9)
    dev = "cuda"
    B = 5
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    cs = (torch.rand(B, C, device=dev) < 0.5).float() / 0.5
    cs2 = (torch.rand(B, C, device=dev) < 0.75).float() / 0.75
    msk = torch.randn(B, H, H, C, device=dev).bfloat16()
    w = (torch.randn(5, C, K, K, device=dev) * 0.03).requires_grad_(True)
    b = torch.randn(5, device=dev).requires_grad_(True)
    Ho = H + 2 * pad - K + 1
    y = torch.zeros(B, 5, Ho, Ho, device=dev)
    wt = torch.empty(K * K * 5 * C, device=dev)
    ops.head_pack(w.detach(), wt)
    ops.head_fwd(x, cs, w.detach(), b.detach(), y, pad, w_t=wt)
    y_generic = torch.zeros_like(y)
    ops.head_fwd(x, cs, w.detach(), b.detach(), y_generic, pad)          # generic kernel (no packed weights)
    assert (y - y_generic).abs().max().item() <= 2e-5
    xin = (x.float().permute(0, 3, 1, 2)).requires_grad_(True)
    ref = torch.sigmoid(F.conv2d(xin * cs[:, :, None, None], w, b, padding=pad))
    assert (y - ref).abs().max().item() <= 2e-5
    dy = torch.randn_like(y)
    dx = torch.zeros_like(x); dx2 = torch.zeros_like(x)
    dw = torch.zeros_like(w); dbias = torch.zeros_like(b)
    ops.head_bwd(x, cs, w.detach(), y, dy, pad, dx, _bits(msk), cs2, 0.2, dx2, dw, dbias, w_t=wt)
    ref.backward(dy)
    gx = xin.grad.permute(0, 2, 3, 1)
    # tensor-core backward (csrc/head_tc.cu): dz and the weights enter the MMAs in bf16 -> bf16 tolerance
    _close(dx, gx, "head dx", rel=6e-3)
    _close(dx2, gx * cs2[:, None, None, :] * torch.where(msk.float() > 0, 1.0, 0.2), "head dx2", rel=8e-3)
    print("head dw rel", rel_err(dw, w.grad), "dbias rel", rel_err(dbias, b.grad))
    assert rel_err(dw, w.grad) <= 4e-3 and rel_err(dbias, b.grad) <= 1e-4
    # ... and against the same gradients computed in fp32 from the bf16-rounded dz / weights: accumulation-order level
    dz = (dy * y * (1 - y)).bfloat16().float()
    xs = (x.float() * cs[:, None, None, :]).permute(0, 3, 1, 2)
    dw_r = torch.nn.grad.conv2d_weight(xs, w.shape, dz, padding=pad)
    dx_r = torch.nn.grad.conv2d_input(xin.shape, w.detach().bfloat16().float(), dz, padding=pad) * cs[:, :, None, None]
    assert rel_err(dw, dw_r) <= 1e-4
    assert rel_err(dx.float(), dx_r.permute(0, 2, 3, 1).bfloat16().float()) <= 2e-3
    ops.head_bwd(x, cs, w.detach(), y, dy, pad, dx, _bits(msk), cs2, 0.2, dx2, dw, dbias, w_t=wt)    # accumulates
    assert rel_err(dw, 2 * dw_r) <= 1e-4


def test_maxpool_fwd_bwd_first_max_tie_rule():
    require_cuda()
    ops = fd().ops
    torch.manual_seed(1)
    dev = "cuda"
    B, H = 3, 30
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    x[:, ::2, ::2] = x[:, 1::2, 1::2]            # many exact ties inside windows
    y = torch.zeros(B, H // 2, H // 2, C, dtype=torch.bfloat16, device=dev)
    ops.maxpool2x2_fwd(x, y)
    xin = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(xin, 2)
    assert torch.equal(y.float(), ref.detach().permute(0, 2, 3, 1))
    gy = torch.randn(B, H // 2, H // 2, C, device=dev).bfloat16()
    msk = torch.randn(B, H, H, C, device=dev).bfloat16()
    cs = (torch.rand(B, C, device=dev) < 0.75).float() / 0.75
    gs = torch.zeros_like(x); gs2 = torch.zeros_like(x)
    ops.maxpool2x2_bwd(x, gy, gs, _bits(msk), cs, 0.2, gs2)
    # the same through the 2-bit positions recorded by the forward (x is not read by the backward)
    amax = torch.empty((B, H // 2, H // 2, C // 8), dtype=torch.int16, device=dev)
    y_b = torch.empty_like(y); ops.maxpool2x2_fwd(x, y_b, amax)
    gs_b = torch.zeros_like(gs); gs2_b = torch.zeros_like(gs2)
    ops.maxpool2x2_bwd(x, gy, gs_b, _bits(msk), cs, 0.2, gs2_b, argmax=amax)
    ref.backward(gy.float().permute(0, 3, 1, 2))
    gref = xin.grad.permute(0, 2, 3, 1)
    assert torch.equal(gs.float(), gref)
    want2 = (gref * cs[:, None, None, :] * torch.where(msk.float() > 0, 1.0, 0.2)).bfloat16()
    assert rel_err(gs2.float(), want2.float()) <= 4e-3
    assert torch.equal(y_b, y) and torch.equal(gs_b, gs) and torch.equal(gs2_b, gs2)


def test_flat_adam_matches_torch_adam():
    """fd_adam_flat on the flat buffer == torch.optim.Adam over the same elements (models/ModelMeta.py:104-112)."""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(3)
    n = 769352
    p0 = torch.randn(n, device="cuda") * 0.1
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(n, device="cuda") * (0.01 * step)
        ref.grad = g.clone()
        opt.step()
        ops.adam_flat(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.0, step)
    assert (p - ref.detach()).abs().max().item() <= 2e-6
    assert rel_err(p - p0, ref.detach() - p0) <= 1e-5


def test_flat_adam_device_step_count_replays_from_a_graph():
    """state != NULL: the kernel keeps the step count / lr on the device; 5 replays of ONE captured launch equal 5
    torch.optim.Adam steps (same gradient buffer, refreshed between replays)."""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(4)
    n = 40000
    p0 = torch.randn(n, device="cuda") * 0.1
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-3)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p); g = torch.zeros_like(p)
    lr_bits = torch.tensor([2e-3], dtype=torch.float32).view(torch.int32).item()
    state = torch.tensor([0, 0, lr_bits, 0], dtype=torch.int32, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            ops.adam_flat(p, g, m, v, 0.0, 0.9, 0.999, 1e-8, 0.0, 0, state)
    torch.cuda.current_stream().wait_stream(side)
    for step in range(1, 6):
        gn = torch.randn(n, device="cuda") * (0.01 * step)
        g.copy_(gn)
        ref.grad = gn.clone()
        opt.step()
        graph.replay()
    torch.cuda.synchronize()
    assert state[:2].tolist() == [5, 0]
    assert (p - ref.detach()).abs().max().item() <= 2e-6
    assert rel_err(p - p0, ref.detach() - p0) <= 1e-5


@pytest.mark.parametrize("u8,B,Hin,Win", [(False, 3, 480, 480), (True, 2, 480, 480), (False, 5, 62, 100), (False, 1, 7, 8)])
def test_stem_stride2_tensor_core_fwd_wgrad(u8, B, Hin, Win):
    """Stem of the standard Resnet (models/Resnet.py:64-70: 3x3, stride 2, pad 1, 3 -> 64) on tcgen05 (csrc/stem_s2_tc.cu):
    forward (bias rides on the ones column of the in-smem im2col tile) and weight / bias gradient against torch fp32.
    The image and the weights are rounded to bf16 on the way into shared memory: bf16 tolerance against the fp32
    reference, fp32-accumulation tolerance against the same convolution on the rounded operands."""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(B * 100 + Win)
    dev = "cuda"
    if u8:
        x = torch.randint(0, 256, (B, 3, Hin, Win), dtype=torch.uint8, device=dev)
        xf = x.float() / 255.0
    else:
        x = torch.rand(B, 3, Hin, Win, device=dev)
        xf = x
    w = (torch.randn(C, 3, 3, 3, device=dev) * 0.2).requires_grad_(True)
    b = torch.randn(C, device=dev).requires_grad_(True)
    Ho, Wo = (Hin - 1) // 2 + 1, (Win - 1) // 2 + 1
    y = torch.full((B, Ho, Wo, C), float("nan"), dtype=torch.bfloat16, device=dev)
    ops.stem_fwd(x, w.detach(), b.detach(), y, 2, 1)
    ref = F.conv2d(xf, w, b, stride=2, padding=1)
    assert tuple(ref.shape) == (B, C, Ho, Wo)
    _close(y, ref.detach().permute(0, 2, 3, 1), "stem s2 fwd", rel=6e-3, mx=2e-2)
    ref_r = F.conv2d(xf.bfloat16().float(), w.detach().bfloat16().float(), b.detach().bfloat16().float(), stride=2, padding=1)
    assert rel_err(y.float(), ref_r.permute(0, 2, 3, 1).bfloat16().float()) <= 2e-3
    g = (torch.randn(B, Ho, Wo, C, device=dev) * 0.1).bfloat16()
    dw = torch.zeros(C, 3, 3, 3, device=dev); dbias = torch.zeros(C, device=dev)
    ops.stem_wgrad(x, g, dw, dbias, 2, 1)
    ref.backward(g.float().permute(0, 3, 1, 2))
    assert rel_err(dw, w.grad) <= 4e-3 and rel_err(dbias, b.grad) <= 1e-4
    wz = torch.zeros_like(w).requires_grad_(True)
    (dref,) = torch.autograd.grad(F.conv2d(xf.bfloat16().float(), wz, None, stride=2, padding=1), wz,
                                  g.float().permute(0, 3, 1, 2))
    assert rel_err(dw, dref) <= 1e-4
    ops.stem_wgrad(x, g, dw, dbias, 2, 1)            # accumulates
    assert rel_err(dw, 2 * dref) <= 1e-4


@pytest.mark.parametrize("nprob,B,H", [(2, 64, 60), (2, 16, 120)])
def test_conv3x3_wgrad_bias_sums_at_full_size(nprob, B, H):
    """Several tiles per CTA at the sizes the bench runs: the bias gradient (column sums of the gradient tiles by the
    epilogue warps) must be complete before the drain's staging tiles overwrite the stage buffers -- a warp that got ahead
    once produced NaN / garbage sums at batch 64 while every small-batch test passed."""
    require_cuda()
    ops = fd().ops
    torch.manual_seed(nprob + B)
    dev = "cuda"
    x = torch.randn(nprob, B, H, H, C, device=dev).bfloat16()
    g = (torch.randn(nprob, B, H, H, C, device=dev) * 0.1).bfloat16()
    n3 = 9 * C * C
    for rep in range(3):
        dw = torch.zeros(nprob, n3, device=dev); db = torch.zeros(nprob, C, device=dev)
        ops.conv3x3_wgrad_multi(x, g, dw.view(-1), n3, db.view(-1), C)
        ref = g.float().sum(dim=(1, 2, 3))
        assert torch.isfinite(db).all() and torch.isfinite(dw).all()
        assert (db - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
