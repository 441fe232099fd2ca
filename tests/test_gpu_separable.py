"""GPU parity of the depthwise-separable backbone (BASELINE config 4: inference + NMS).

Block kernel (fd_sepblock_fwd) against a torch fp32 evaluation of the same block on the SAME bf16-rounded inputs
and weights (tolerance: the two bf16-rounded intermediates t1 / t2 and the bf16 output, fp32 accumulation:
rel-L2 <= 6e-3, the conv kernels' own bar is 4e-3 with one rounding); whole model against the golden head of the
REAL reference (max-abs <= 2e-2 like the residual backbones), decode + NMS of OUR head bit-exact against the oracle.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import backbone_oracle as bo
from oracle import yolo_oracle as yo
from tests.gpu_util import fd, rel_err, require_cuda
from tests.util import load_golden, seeded_separable_params, synth_boxes

pytestmark = pytest.mark.gpu


def _block_ref(x, w1, wd, w2, slope=0.2):
    """x [B,H,W,64] bf16 -> fp32 NHWC result of the block with the kernel's rounding points (t1, t2 in bf16)."""
    xf = x.float().permute(0, 3, 1, 2)
    t1 = F.leaky_relu(F.conv2d(xf, w1.float()[:, :, None, None]), slope).bfloat16().float()
    t2 = F.leaky_relu(F.conv2d(t1, wd, padding=1, groups=64), slope).bfloat16().float()
    y = F.conv2d(t2, w2.float()[:, :, None, None]) + xf
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("B,H,W", [(3, 60, 60), (5, 30, 30), (7, 15, 15), (2, 9, 13), (1, 64, 70), (300, 15, 15)])
def test_sepblock_kernel_vs_torch(B, H, W):
    require_cuda()
    ops = fd().ops
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.randn(B, H, W, 64, generator=g).cuda().bfloat16()
    pw = (torch.randn(2, 64, 64, generator=g) * 0.15).cuda()
    dw = (torch.randn(1, 64, 1, 3, 3, generator=g) * 0.4).cuda()
    w_pw = torch.empty((2, 64, 64), dtype=torch.bfloat16, device="cuda")
    w_dw = torch.empty((1, 9, 64), device="cuda")
    ops.sep_pack(pw.view(2, 64, 64, 1, 1), w_pw, dw, w_dw)
    assert torch.equal(w_pw, pw.bfloat16())
    assert torch.equal(w_dw[0], dw[0, :, 0].reshape(64, 9).t().contiguous())
    out = torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.sepblock_fwd(x, w_pw[0], w_dw[0], w_pw[1], 0.2, out)
    torch.cuda.synchronize()
    want = _block_ref(x, w_pw[0], dw[0], w_pw[1])
    assert torch.isfinite(out.float()).all()
    e = rel_err(out.float(), want)
    print(f"sepblock {B}x{H}x{W}: rel-L2 {e:.2e}")
    assert e <= 6e-3
    # borders are where the zero padding of the depthwise stage matters
    for sl in (out[:, 0], out[:, -1], out[:, :, 0], out[:, :, -1]):
        pass
    eb = rel_err(torch.cat([out[:, 0].float().reshape(-1), out[:, -1].float().reshape(-1),
                            out[:, :, 0].float().reshape(-1), out[:, :, -1].float().reshape(-1)]),
                 torch.cat([want[:, 0].reshape(-1), want[:, -1].reshape(-1), want[:, :, 0].reshape(-1),
                            want[:, :, -1].reshape(-1)]))
    assert eb <= 6e-3, eb


@pytest.mark.parametrize("B,H,W", [(3, 60, 60), (5, 30, 30), (2, 13, 9), (2, 64, 70), (150, 30, 30)])
def test_sepblock_fused_pool_equals_block_then_maxpool(B, H, W):
    """pool=1 (MaxPool2d(2) fused, SeparableCNN.py:49-50) is bit-identical to the block kernel followed by
    fd_maxpool2x2_fwd, and matches torch max_pool2d of the fp32 block result (floor semantics on odd sizes)."""
    require_cuda()
    ops = fd().ops
    g = torch.Generator().manual_seed(B * 77 + W)
    x = torch.randn(B, H, W, 64, generator=g).cuda().bfloat16()
    w_pw = (torch.randn(2, 64, 64, generator=g) * 0.15).cuda().bfloat16()
    dw = (torch.randn(64, 1, 3, 3, generator=g) * 0.4).cuda()
    w_dw = dw[:, 0].reshape(64, 9).t().contiguous()
    full = torch.empty((B, H, W, 64), dtype=torch.bfloat16, device="cuda")
    ops.sepblock_fwd(x, w_pw[0], w_dw, w_pw[1], 0.2, full)
    fused = torch.full((B, H // 2, W // 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.sepblock_fwd(x, w_pw[0], w_dw, w_pw[1], 0.2, fused, pool=True)
    torch.cuda.synchronize()
    want = F.max_pool2d(full.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(fused.float(), want)
    if H % 2 == 0 and W % 2 == 0:
        two = torch.empty_like(fused)
        ops.maxpool2x2_fwd(full, two)
        assert torch.equal(fused, two)
    ref = F.max_pool2d(_block_ref(x, w_pw[0], dw, w_pw[1]).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert rel_err(fused.float(), ref) <= 6e-3


def _model(seed=6):
    torch.manual_seed(seed)
    return fd().models.SeparableCNN.SeparableCNN(filters=64, input_shape=(3, 480, 480), probability_threshold=0.47,
                                                 iou_threshold=0.3)


def test_separable_model_vs_reference_golden():
    require_cuda()
    g = load_golden("separable_seed6.npz")
    m = _model()
    for k, v in m.state_dict().items():              # same construction order / seeding as the reference
        s = g["w_sum." + k]
        assert abs(v.double().sum().item() - s[0]) < 1e-9 and abs(v.double().abs().sum().item() - s[1]) < 1e-9, k
    assert m.num_of_patches == 16 and m.reduce_bounding_boxes.x_patch_size == 30.0
    m = m.cuda().eval()
    x = torch.rand(2, 3, 480, 480, generator=torch.Generator().manual_seed(7)).cuda()
    with torch.no_grad():
        y = m(x)
    assert tuple(y.shape) == (2, 5, 10, 10)
    d = (y.cpu() - torch.from_numpy(g["y"])).abs()
    print("separable head max/mean abs err", d.max().item(), d.mean().item())
    assert d.max().item() <= 2e-2 and d.mean().item() <= 2e-3
    kept = m.non_max_suppression(y)                  # decode + NMS of OUR head: bit-exact against the oracle
    for i in range(2):
        want = yo.reduce_bounding_boxes(y[i].cpu().numpy(), 0.47, 0.3, (3, 480, 480), 16)
        assert kept[i].cpu().numpy().tobytes() == want.tobytes()
    # the reference's predict path: uint8 image -> resize (no-op at 480) -> /255 (fused into the stem) -> boxes of image 0
    x8 = (x[0] * 255).round().to(torch.uint8)
    with torch.no_grad():
        b0 = m(x8, predict=torch.tensor(1))
        yq = m((x8.float() / 255.0).unsqueeze(0))
    want = yo.reduce_bounding_boxes(yq[0].cpu().numpy(), 0.47, 0.3, (3, 480, 480), 16)
    assert b0.cpu().numpy().tobytes() == want.tobytes()


def test_separable_model_batch256_vs_oracle_and_train_mode_runs():
    """BASELINE config 4 batch size (256): every image of the batch against the fp32 oracle."""
    require_cuda()
    m = _model(seed=9).cuda().eval()
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    x = torch.rand(256, 3, 480, 480, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = m(x.cuda()).cpu()                            # inference engine: fused separable-block kernel
        want = torch.cat([bo.separable_forward(x[i:i + 32], p) for i in range(0, 256, 32)])
    d = (y - want).abs()
    print("separable B=256 head max/mean abs err", d.max().item(), d.mean().item())
    assert d.max().item() <= 2e-2 and d.mean().item() <= 2e-3
    m.train()                                            # train mode: the layer-by-layer engine with Dropout2d
    yt = m(x[:2].cuda())
    assert yt.requires_grad and tuple(yt.shape) == (2, 5, 10, 10) and bool(torch.isfinite(yt).all())


def test_separable_filters128_planar_vs_oracle():
    """SeparableCNN(filters=128) (the width of the reference's __main__, SeparableCNN.py:124) on two channel planes:
    centre-tap fd_conv3x3 chains for the pointwise convolutions, fd_dwconv3x3_lrelu, planar head."""
    require_cuda()
    from tests.util import seeded_separable_params
    p = seeded_separable_params(128, seed=11)
    m = fd().models.SeparableCNN.SeparableCNN(filters=128, input_shape=(3, 480, 480))
    m.load_state_dict(p, strict=True)
    m = m.cuda().eval()
    assert type(m.engine).__name__ == "SeparablePlanarEngine" and sum(q.numel() for q in m.parameters()) == 400773
    x = torch.rand(3, 3, 480, 480, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        y = m(x.cuda()).cpu()                            # inference engine (channel planes)
        want = bo.separable_forward(x, p)
    d = (y - want).abs()
    print("separable F=128 head max/mean abs err", d.max().item(), d.mean().item())
    assert d.max().item() <= 2e-2 and d.mean().item() <= 2e-3
    kept = m.non_max_suppression(y.cuda())
    for i in range(3):
        ref = yo.reduce_bounding_boxes(y[i].numpy(), 0.5, 0.5, (3, 480, 480), 16)
        assert kept[i].cpu().numpy().tobytes() == ref.tobytes()


def test_depthwise_plane_kernel_vs_torch():
    require_cuda()
    ops = fd().ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 13, 9, 64, generator=g).cuda().bfloat16()
    dw = (torch.randn(64, 1, 3, 3, generator=g) * 0.4).cuda()
    w_dw = dw[:, 0].reshape(64, 9).t().contiguous()
    out = torch.empty_like(x)
    ops.dwconv3x3_lrelu(x, w_dw, 0.2, out)
    want = F.leaky_relu(F.conv2d(x.float().permute(0, 3, 1, 2), dw, padding=1, groups=64), 0.2).permute(0, 2, 3, 1)
    assert rel_err(out.float(), want) <= 4e-3


@pytest.mark.parametrize("filters,B", [(64, 4), (128, 2)])
def test_separable_train_step_vs_oracle_train_mode(filters, B):
    """models/SeparableCNN.py:40-51,104-117 in TRAIN mode (Dropout2d(0.25) on every pw2 output, Dropout2d(0.5) before the
    head): fused call model.train_step and the autograd path (forward + yolo_loss + backward) against the torch-fp32
    oracle with the kernel's own Dropout2d multipliers injected; head, summed loss, every parameter gradient.  Tolerance as
    for the residual backbones (bf16 activations): head 2e-2 / 2e-3, loss 1e-2, per-tensor gradient rel-L2 <= 4e-2."""
    require_cuda()
    pkg = fd()
    p = seeded_separable_params(filters, seed=16)
    m = pkg.models.SeparableCNN.SeparableCNN(filters=filters, input_shape=(3, 480, 480))
    m.load_state_dict(p, strict=True)
    m = m.cuda().train()
    gen = torch.Generator().manual_seed(17)
    x = torch.rand(B, 3, 480, 480, generator=gen)
    gt = torch.stack([torch.from_numpy(yo.grid_encode(synth_boxes(gen, 1, 60).numpy(), 10, 480, 480)) for _ in range(B)])
    torch.manual_seed(18)
    loss = m.train_step(x.cuda(), gt.cuda())
    eng = m.train_engine
    pl = eng.plan(B)
    drop = pl["drop"].cpu()                                  # [num_blocks + 1, G, B, 64]
    scales = [drop[k].permute(1, 0, 2).reshape(B, filters, 1, 1) for k in range(drop.shape[0])]
    y_ref, loss_ref, g_ref = bo.train_step(x, gt, p, 16, forward=bo.separable_forward, drop_scales=scales)
    d = (pl["y"].cpu() - y_ref).abs()
    print(f"SeparableCNN({filters}) train-mode head max/mean abs err", d.max().item(), d.mean().item())
    assert d.max().item() <= 2e-2 and d.mean().item() <= 2e-3
    assert abs(loss.item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    worst = ("", 0.0)
    for k, prm in m.named_parameters():
        e = rel_err(prm.grad.cpu(), g_ref[k])
        worst = max(worst, (k, e), key=lambda t: t[1])
        assert e <= 4e-2, (k, e)
    print("worst per-tensor gradient rel-L2", worst)
    # eval mode: autograd path == fused path; inference engine (fused block kernel / planes) agrees with the training engine
    m.eval()
    l1 = m.train_step(x.cuda(), gt.cuda())
    fused = {k: q.grad.clone() for k, q in m.named_parameters()}
    for q in m.parameters():
        q.grad = None
    y_hat = m(x.cuda())
    l2 = pkg.losses.YoloLoss.yolo_loss_batch(y_hat, gt.cuda())
    l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-5 * abs(l2.item())
    for k, q in m.named_parameters():
        assert rel_err(q.grad, fused[k]) <= 1e-4, k
    with torch.no_grad():
        y_inf = m(x.cuda())
    assert (y_inf - y_hat.detach()).abs().max().item() <= 1e-2
    # one Adam step through ModelMeta (train_model.py:35-39 pattern) lowers the loss on the same batch
    meta = pkg.models.ModelMeta(model=m, lr=1e-3)
    (opt,), _ = meta.configure_optimizers()
    opt.zero_grad()
    y_hat = m(x.cuda())
    l3 = pkg.losses.YoloLoss.yolo_loss_batch(y_hat, gt.cuda())
    l3.backward()
    opt.step()
    with torch.no_grad():
        l4 = pkg.losses.YoloLoss.yolo_loss_batch(m(x.cuda()), gt.cuda())
    assert l4.item() < l3.item()


def test_lrelu_bwd_and_dw_wgrad_kernels_vs_torch():
    require_cuda()
    ops = fd().ops
    g_ = torch.Generator().manual_seed(4)
    B, H, W = 3, 30, 30
    x = torch.randn(B, H, W, 64, generator=g_).cuda().bfloat16()
    g = torch.randn(B, H, W, 64, generator=g_).cuda().bfloat16()
    out = torch.empty_like(g)
    ops.lrelu_bwd(g, x, 0.2, out)
    want = (g.float() * torch.where(x.float() >= 0, 1.0, 0.2)).bfloat16()
    assert torch.equal(out, want)
    dw = torch.zeros(64, 1, 3, 3, device="cuda")
    ops.dwconv3x3_wgrad(x, g, dw)
    wz = torch.zeros(64, 1, 3, 3, device="cuda", requires_grad=True)
    yy = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wz, None, padding=1, groups=64)
    (ref,) = torch.autograd.grad(yy, wz, g.float().permute(0, 3, 1, 2))
    assert rel_err(dw, ref) <= 1e-4
