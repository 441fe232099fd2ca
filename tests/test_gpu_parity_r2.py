"""Round-2 parity tests of the configurations that are actually benchmarked and shipped:

* the TRAIN-MODE step (Dropout2d active) at the bench batch size against the oracle, with the kernel's own Dropout2d
  multipliers injected into the oracle (``drop_scales``) -- PoolResnet B = 64 (bench.py's workload) and Resnet B = 16;
* all four official checkpoints over ALL 24 ``imgs/test_imgs`` frames through the demo path (demo_model.py:16-37);
* ``BaseModel.predict`` including the GPU resize (``fd_resize_bilinear``);
* the stale-backward guard (two forwards, one backward).

Stated tolerance (bf16 activations, fp32 accumulation): head max-abs <= 2e-2 / mean-abs <= 2e-3 (MobilenetV3, ~50 bf16
layer boundaries incl. SqueezeExcite gates and linear bottlenecks: 1.5e-1 / 6e-3 -- a CPU emulation of the same bf16
rounding points gives 8.9e-2 / 2.3e-3 over the 24 frames, the GPU 9.4e-2 / 2.7e-3), summed loss rel <= 1e-2, per-tensor
gradient rel-L2 <= GRAD_REL.
"""
import importlib

import cv2
import numpy as np
import pytest
import torch

from oracle import backbone_oracle as bo
from oracle import yolo_oracle as yo
from tests.gpu_util import fd, rel_err, require_cuda
from tests.util import load_golden, synth_boxes

pytestmark = pytest.mark.gpu

HEAD_MAX, HEAD_MEAN, LOSS_REL, GRAD_REL = 2e-2, 2e-3, 1e-2, 2e-2       # GRAD_REL = 2x the measured worst case (9.7e-3)


def _frames():
    im = load_golden("official_images.npz")
    out = []
    for i in range(len(im["names"])):
        rgb = cv2.imdecode(im["png"][im["offsets"][i]:im["offsets"][i + 1]], cv2.IMREAD_UNCHANGED)
        out.append(torch.from_numpy(rgb).permute(2, 0, 1).contiguous())
    return out


@pytest.mark.parametrize("arch,B,S,kmin,kmax", [("PoolResnet", 64, 10, 1, 100), ("Resnet", 16, 15, 101, 400)])
def test_train_mode_step_vs_oracle_with_injected_dropout(arch, B, S, kmin, kmax):
    """model.train(): Dropout2d(0.25) in every block and Dropout2d(0.5) before the head (PoolResnet.py:31,39,69,100).
    The multipliers the kernels used are read back from the plan and handed to the oracle, so the comparison covers
    the whole mask wiring (forward epilogues, dgrad chain, un-pool, head backward) at the benchmarked batch size."""
    require_cuda()
    pkg = fd()
    Model = getattr(getattr(pkg.models, arch), arch)
    torch.manual_seed(31)
    m = Model(filters=64, input_shape=(3, 480, 480), num_of_patches=S).cuda().train()
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(32)
    x = torch.rand(B, 3, 480, 480, generator=gen)
    gt = torch.stack([torch.from_numpy(yo.grid_encode(synth_boxes(gen, kmin, kmax).numpy(), S, 480, 480)) for _ in range(B)])
    torch.manual_seed(33)
    loss = m.train_step(x.cuda(), gt.cuda())
    pl = m.engine.plan(B, True)
    assert pl.drop is not None
    drop = pl.drop.cpu()                                    # [num_blocks + 1, B, 64]
    kept = (drop[:-1] > 0).float().mean().item()
    vals = drop[:-1].unique().tolist()                                                        # Dropout2d(0.25)
    assert 0.70 < kept < 0.80 and all(v == 0.0 or abs(v - 1.0 / 0.75) < 1e-6 for v in vals), vals
    assert set(drop[-1].unique().tolist()) <= {0.0, 2.0}                                      # Dropout2d(0.5)
    scales = [drop[k].view(B, 64, 1, 1) for k in range(drop.shape[0])]
    fwd = bo.poolresnet_forward if arch == "PoolResnet" else bo.resnet_forward
    y_ref, loss_ref, g_ref = bo.train_step(x, gt, p, S, forward=fwd, drop_scales=scales)
    d = (pl.y.cpu() - y_ref).abs()
    print(f"{arch} train-mode B={B}: head max/mean abs err", d.max().item(), d.mean().item())
    assert d.max().item() <= HEAD_MAX and d.mean().item() <= HEAD_MEAN
    print("loss", loss.item(), loss_ref.item())
    assert abs(loss.item() - loss_ref.item()) <= LOSS_REL * abs(loss_ref.item())
    worst = ("", 0.0)
    for k, prm in m.named_parameters():
        e = rel_err(prm.grad.cpu(), g_ref[k])
        if e > worst[1]:
            worst = (k, e)
        assert e <= GRAD_REL, (k, e)
    print("worst per-tensor gradient rel-L2", worst)


def _box_check(name, heads_ours, boxes_ours, g, S_dec, head_max, head_mean):
    """Heads within tolerance; and wherever OUR head selects the same candidate cells as the golden head (conf > thr),
    the demo path must return the same number of boxes with coordinates within head-tolerance * 480 + rounding."""
    p_thr = float(g["p_thr"])
    n_img = heads_ours.shape[0]
    d = (heads_ours - torch.from_numpy(g["heads"])).abs()
    print(f"{name}: head max/mean abs err over {n_img} frames", d.max().item(), d.mean().item())
    assert d.max().item() <= head_max and d.mean().item() <= head_mean
    same_cand = same_count = 0
    for i in range(n_img):
        cand_ref = g["heads"][i, 0] > p_thr
        cand_our = heads_ours[i, 0].numpy() > p_thr
        want = g["boxes"][i, :g["counts"][i]]
        got = boxes_ours[i]
        if not np.array_equal(cand_ref, cand_our):
            continue
        same_cand += 1
        if got.shape[0] != want.shape[0]:
            continue            # an NMS decision with IoU at the 0.01 threshold flipped under the head tolerance
        same_count += 1
        if want.shape[0]:
            order_w = np.lexsort((want[:, 2], want[:, 1]))
            order_g = np.lexsort((got[:, 2], got[:, 1]))
            tol_px = d[i].max().item() * 480.0 + 1.0
            assert np.abs(got[order_g][:, 0] - want[order_w][:, 0]).max() <= head_max
            assert np.abs(got[order_g][:, 1:] - want[order_w][:, 1:]).max() <= tol_px, (name, i)
    print(f"{name}: identical candidate cells on {same_cand}/{n_img} frames, identical box count on {same_count} of those")
    assert same_cand >= n_img // 2 and same_count >= same_cand - max(1, same_cand // 8)


def _run_demo_path(m, frames):
    heads, boxes = [], []
    for t in frames:
        t2 = torch.stack([t, t]).cuda()                          # demo_model.py:20
        with torch.no_grad():
            h = m(t2.float() / 255.0)
            b = m(t2, predict=torch.tensor(1))                   # demo_model.py:21 (uint8 in, /255 fused into the stem)
        heads.append(h[0].cpu())
        boxes.append(b.cpu().numpy().reshape(-1, 5))
        # decode + NMS of OUR head is bit-exact against the oracle on the same head
        rb = m.reduce_bounding_boxes
        with torch.no_grad():
            h8 = m(t2)[0]                                         # uint8 path head (what predict=1 decodes)
        want = yo.reduce_bounding_boxes(h8.cpu().numpy(), rb.probability_threshold, rb.iou_threshold, (3, 480, 480),
                                        rb.num_of_patches)
        assert boxes[-1].astype(np.float32).tobytes() == want.tobytes()
    return torch.stack(heads), boxes


@pytest.mark.parametrize("arch,golden,filters,S", [("PoolResnet", "official_poolresnet_medium.npz", 64, 10),
                                                   ("PoolResnet", "official_poolresnet_small.npz", 32, 10),
                                                   ("Resnet", "official_resnet_medium.npz", 64, 15)])
def test_official_checkpoints_all_24_frames(arch, golden, filters, S):
    require_cuda()
    pkg = fd()
    g = load_golden(golden)
    src = load_golden("official_medium.npz") if golden == "official_poolresnet_medium.npz" else g
    sd = {k[3:]: torch.from_numpy(src[k]) for k in src.files if k.startswith("sd.")}
    Model = getattr(getattr(pkg.models, arch), arch)
    m = Model(filters=filters, input_shape=(3, 480, 480), num_of_patches=S, probability_threshold=float(g["p_thr"]),
              iou_threshold=float(g["iou_thr"]))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    heads, boxes = _run_demo_path(m, _frames())
    _box_check(f"{arch}-{filters}", heads, boxes, g, S, HEAD_MAX, HEAD_MEAN)


def test_small_checkpoint_train_step_on_padded_planes():
    """filters = 32 runs on the 64-channel engine with zero-padded weights (PaddedBackboneEngine): train step against
    the oracle -- the gradients of the logical channels are unchanged, those of the padding exactly zero."""
    require_cuda()
    pkg = fd()
    g = load_golden("official_poolresnet_small.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    m = pkg.models.PoolResnet.PoolResnet(filters=32, input_shape=(3, 480, 480), num_of_patches=10)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    assert type(m.engine).__name__ == "PaddedBackboneEngine" and sum(q.numel() for q in m.parameters()) == 200357
    B = 3
    gen = torch.Generator().manual_seed(41)
    x = torch.rand(B, 3, 480, 480, generator=gen)
    gt = torch.stack([torch.from_numpy(yo.grid_encode(synth_boxes(gen, 1, 100).numpy(), 10, 480, 480)) for _ in range(B)])
    y_ref, loss_ref, g_ref = bo.train_step(x, gt, sd, 10)
    loss = m.train_step(x.cuda(), gt.cuda())
    assert abs(loss.item() - loss_ref.item()) <= LOSS_REL * abs(loss_ref.item())
    # a TRAINED checkpoint at batch 3: gradients are sums of nearly cancelling terms, bf16 noise weighs more than at
    # initialisation (measured up to 1.1e-1 on single tensors) -- per tensor <= 2e-1, all tensors together <= 8e-2
    num = den = 0.0
    worst = ("", 0.0)
    for k, prm in m.named_parameters():
        assert prm.grad.shape == g_ref[k].shape
        e = rel_err(prm.grad.cpu(), g_ref[k])
        worst = max(worst, (k, e), key=lambda t: t[1])
        assert e <= 2e-1, (k, e)
        num += (prm.grad.cpu().double() - g_ref[k].double()).pow(2).sum().item()
        den += g_ref[k].double().pow(2).sum().item()
    print("small checkpoint: worst per-tensor gradient rel-L2", worst, "global", (num / den) ** 0.5)
    assert (num / den) ** 0.5 <= 8e-2
    eng = m.engine
    big = eng.gflat.clone()
    big[eng.index.long()] = 0
    assert float(big.abs().max()) == 0.0                 # gradients of the padded channels are exactly zero
    opt = m.flat_optimizer(lr=1e-3)
    before = m.conv1.weight.detach().clone()
    m.train_step(x.cuda(), gt.cuda(), optimizer=opt)
    assert (m.conv1.weight.detach() - before).abs().max().item() > 1e-4      # nn.Parameters see the flat Adam update


def test_mobilenetv3_official_checkpoint_all_24_frames():
    """BASELINE config 4: MobilenetV3Backbone (timm tf_mobilenetv3_small_100 graph recovered from the archive) with the
    official weights over the 24 frames: raw [5,15,15] heads against the archive's, boxes of the demo path."""
    require_cuda()
    pkg = fd()
    g = load_golden("official_mobilenetv3.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    m = pkg.models.MobilenetV3Backbone.MobilenetV3Backbone(576, (3, 480, 480), 15, probability_threshold=float(g["p_thr"]),
                                                           iou_threshold=float(g["iou_thr"]))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    frames = _frames()
    heads, boxes = _run_demo_path(m, frames)
    d = (heads - torch.from_numpy(g["heads"])).abs()
    print("MobilenetV3: head max/mean abs err over 24 frames", d.max().item(), d.mean().item())
    assert d.max().item() <= 1.5e-1 and d.mean().item() <= 6e-3
    assert (d > 5e-2).float().mean().item() <= 1e-3                  # the tail: < 0.1 % of the 27 000 head values
    # the whole batch in one call == frame by frame (batch-size independent kernels)
    xb = torch.stack(frames).cuda()
    with torch.no_grad():
        hb = m(xb.float() / 255.0).cpu()
    assert torch.equal(hb, heads)                                    # deterministic kernels: bit-identical
    # against the oracle restatement on seeded random images, odd batch size
    x = torch.rand(5, 3, 480, 480, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        y_ref = bo.mobilenetv3_forward(x, sd)
    e = (y - y_ref).abs()
    print("MobilenetV3 vs oracle on random images: max/mean", e.max().item(), e.mean().item())
    assert e.max().item() <= 1.5e-1 and e.mean().item() <= 6e-3
    kept = m.non_max_suppression(y.cuda())
    for i in range(5):
        want = yo.reduce_bounding_boxes(y[i].numpy(), float(g["p_thr"]), float(g["iou_thr"]), (3, 480, 480), 15)
        assert kept[i].cpu().numpy().tobytes() == want.tobytes()


@pytest.mark.parametrize("M,K,N,act,res", [(1000, 16, 72, 1, False), (257, 24, 88, 1, False), (900, 88, 24, 0, True),
                                           (3600, 96, 40, 0, False), (1800, 40, 240, 2, False), (450, 240, 40, 0, True),
                                           (333, 288, 96, 0, False), (450, 576, 96, 0, True), (225, 96, 576, 2, False),
                                           (130, 48, 288, 2, False), (14400 * 3, 16, 16, 0, False), (77, 144, 48, 0, True)])
def test_pw_conv_vs_torch(M, K, N, act, res):
    """fd_pw_conv (tcgen05 GEMM, channel counts of the MobilenetV3 graph incl. partial K slabs, split N, ragged M) against
    torch fp32 on the same bf16-rounded operands."""
    require_cuda()
    ops = fd().ops
    g = torch.Generator().manual_seed(M + K + N)
    x = (torch.randn(M, K, generator=g)).cuda().bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    r = (torch.randn(M, N, generator=g)).cuda().bfloat16() if res else None
    packed = torch.empty(ops.pw_packed_elems(N, K), dtype=torch.bfloat16, device="cuda")
    ops.pw_pack(w, scale, packed)
    bpad = torch.zeros(ops.pw_padded_n(N), device="cuda")
    bpad[:N] = bias
    out = torch.full((M, N), 7.0, device="cuda").bfloat16()
    ops.pw_conv(x, packed, bpad, N, act, out, residual=r)
    wq = (w * scale[:, None]).bfloat16().float()
    ref = x.float() @ wq.t() + bias
    ref = torch.relu(ref) if act == 1 else torch.nn.functional.hardswish(ref) if act == 2 else ref
    if res:
        ref = ref + r.float()
    err = (out.float() - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item())
    assert err <= tol, (err, tol)
    assert rel_err(out.float(), ref) <= 6e-3


@pytest.mark.parametrize("C,K,s,H,W,act,se", [(16, 3, 2, 240, 240, 1, True), (72, 3, 2, 120, 120, 1, False),
                                               (88, 3, 1, 60, 60, 1, False), (96, 5, 2, 60, 60, 2, True),
                                               (240, 5, 1, 30, 30, 2, True), (288, 5, 2, 30, 30, 2, True),
                                               (576, 5, 1, 15, 15, 2, True), (24, 5, 2, 33, 47, 2, True)])
def test_dwconv_se_vs_torch(C, K, s, H, W, act, se):
    """fd_dwconv (TF-SAME asymmetric padding for stride 2, fused bias + activation + SE channel sums), fd_se_gate and
    fd_scale_channels against torch fp32."""
    require_cuda()
    ops = fd().ops
    B = 3
    g = torch.Generator().manual_seed(C * K + H)
    x = torch.randn(B, C, H, W, generator=g).bfloat16()
    w = torch.randn(C, 1, K, K, generator=g) / K
    scale = torch.rand(C, generator=g) + 0.5
    bias = torch.randn(C, generator=g) * 0.1
    wq = w * scale.view(-1, 1, 1, 1)
    ref = bo._conv_same(x.float(), wq, s, groups=C) + bias.view(1, -1, 1, 1)
    ref = torch.relu(ref) if act == 1 else torch.nn.functional.hardswish(ref) if act == 2 else ref
    Ho, Wo = ref.shape[-2:]
    pad = (bo._same_pad(H, K, s)[0], bo._same_pad(W, K, s)[0]) if s > 1 else (K // 2, K // 2)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = torch.empty(K * K, C, device="cuda")
    ops.dw_pack(w.cuda().contiguous(), scale.cuda(), wp)
    out = torch.empty(B, Ho, Wo, C, dtype=torch.bfloat16, device="cuda")
    ssum = torch.full((B, ops.dwconv_se_blocks(Ho, Wo, C), C), 3.0, device="cuda") if se else None
    ops.dwconv(xd, wp, bias.cuda(), K, s, pad[0], pad[1], act, out, se_partial=ssum)
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    if se:
        want_sum = got.sum((2, 3))
        assert rel_err(ssum.sum(1).cpu(), want_sum) <= 1e-4
        ssum2 = torch.empty_like(ssum)
        ops.dwconv(xd, wp, bias.cuda(), K, s, pad[0], pad[1], act, torch.empty_like(out), se_partial=ssum2)
        assert torch.equal(ssum, ssum2)                                 # deterministic (no atomics)
        R = max(8, C // 4)
        w1, b1 = torch.randn(R, C, generator=g) / C ** 0.5, torch.randn(R, generator=g) * 0.1
        w2, b2 = torch.randn(C, R, generator=g) / R ** 0.5, torch.randn(C, generator=g) * 0.1
        gate = torch.empty(B, C, device="cuda")
        ops.se_gate(ssum, Ho * Wo, w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), gate)
        mean = want_sum / (Ho * Wo)
        gate_ref = torch.nn.functional.hardsigmoid(torch.relu(mean @ w1.t() + b1) @ w2.t() + b2)
        assert (gate.cpu() - gate_ref).abs().max().item() <= 1e-4
        ops.scale_channels(out, gate)
        want = (got * gate_ref.view(B, C, 1, 1)).bfloat16().float()
        assert (out.float().permute(0, 3, 1, 2).cpu() - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("h,w", [(344, 450), (898, 1600), (480, 640), (120, 97)])
def test_resize_bilinear_vs_torch(dtype, h, w):
    """fd_resize_bilinear == F.interpolate(bilinear, align_corners=False) of the pinned torchvision Resize: float within
    1e-4 (of 255), uint8 identical except where the fp32 value sits on a rounding tie (<= 1 LSB, < 0.1 % of pixels)."""
    require_cuda()
    pkg = fd()
    g = torch.Generator().manual_seed(h + w)
    x = torch.randint(0, 256, (2, 3, h, w), generator=g, dtype=torch.uint8)
    if dtype == torch.float32:
        x = x.float() + torch.rand(2, 3, h, w, generator=g)
    want = bo.resize_bilinear(x, (480, 480))
    resize_to = importlib.import_module(pkg.__name__ + ".models.BaseModel").resize_to
    got = resize_to(x.cuda(), (480, 480)).cpu()
    assert got.dtype == want.dtype and got.shape == want.shape
    diff = (got.float() - want.float()).abs()
    if dtype == torch.uint8:
        assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3
    else:
        assert diff.max().item() <= 1e-2            # values up to 256: fp32 rounding of the interpolation weights
    one = resize_to(x[0].cuda(), (480, 480)).cpu()      # 3-D input
    assert torch.equal(one, got[0])


def test_basemodel_predict_with_gpu_resize():
    """models/BaseModel.py:56-71 ``predict``: Resize -> /255 -> forward -> NMS; returns (image, boxes of image 0).  A host
    image of another size goes through fd_resize_bilinear; checked against the oracle chain."""
    require_cuda()
    pkg = fd()
    src = load_golden("official_medium.npz")
    sd = {k[3:]: torch.from_numpy(src[k]) for k in src.files if k.startswith("sd.")}
    m = pkg.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    frame = _frames()[0]
    small = bo.resize_bilinear(frame, (360, 500))                     # a non-480 uint8 image
    for x in (small, small.float()):
        image, boxes = m.predict(x, probability_threshold=0.6, iou_threshold=0.3)
        assert m.reduce_bounding_boxes.probability_threshold == 0.6 and m.reduce_bounding_boxes.iou_threshold == 0.3
        x480 = bo.resize_bilinear(x, (480, 480))
        img_ref = x480 / 255.0
        assert tuple(image.shape) == (3, 480, 480) and (image.cpu() - img_ref).abs().max().item() <= 1.01 / 255
        with torch.no_grad():
            y_ref = bo.poolresnet_forward(img_ref.unsqueeze(0), sd, 10)
            y = m(image.unsqueeze(0))
        assert (y.cpu() - y_ref).abs().max().item() <= HEAD_MAX
        want = yo.reduce_bounding_boxes(y[0].cpu().numpy(), 0.6, 0.3, (3, 480, 480), 10)
        assert boxes.cpu().numpy().reshape(-1, 5).astype(np.float32).tobytes() == want.tobytes()


def test_backward_after_second_forward_raises():
    """The saved activations live in the engine's per-batch-size plan: a second forward of the same batch size before
    backward() overwrites them -- that must raise, not silently mix two batches."""
    require_cuda()
    pkg = fd()
    torch.manual_seed(0)
    m = pkg.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10).cuda().eval()
    x1, x2 = torch.rand(2, 3, 480, 480).cuda(), torch.rand(2, 3, 480, 480).cuda()
    y1 = m(x1)
    _ = m(x2)
    with pytest.raises(RuntimeError, match="ANOTHER forward"):
        y1.sum().backward()
    y3 = m(x1)
    with torch.no_grad():
        _ = m(x2)                         # eval-plan forward: does not touch the training plan
    y3.sum().backward()
    assert all(p.grad is not None for p in m.parameters())


def test_torchscript_archive_runs_demo_extract_face(tmp_path):
    """the official "medium" weights re-exported with to_torchscript, re-loaded with torch.jit.load and
    driven exactly like demo_model.py:16-21 (HOST uint8 [2,3,480,480], predict=torch.tensor(1)): identical rows to the
    eager predict path, returned on the host like the reference's CPU archive."""
    require_cuda()
    pkg = fd()
    src = load_golden("official_medium.npz")
    g = load_golden("official_poolresnet_medium.npz")
    sd = {k[3:]: torch.from_numpy(src[k]) for k in src.files if k.startswith("sd.")}
    m = pkg.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10,
                                         probability_threshold=float(g["p_thr"]), iou_threshold=float(g["iou_thr"]))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    path = tmp_path / "medium_b200.pth"
    m.to_torchscript(path)
    ts = torch.jit.load(str(path))                                   # demo_model.py:11-13
    frames = _frames()
    n_boxes = 0
    for i in (0, 2, 15, 19):
        t2 = torch.stack([frames[i], frames[i]])                     # host tensor, demo_model.py:19-20
        with torch.no_grad():
            b_ts = ts(t2, predict=torch.tensor(1))                   # demo_model.py:21
            b_eager = m(t2.cuda(), predict=torch.tensor(1))
            head = ts(t2.cuda().float() / 255.0)
        assert not b_ts.is_cuda and torch.equal(b_ts, b_eager.cpu())
        assert tuple(head.shape) == (2, 5, 10, 10) and head.is_cuda
        for b in b_ts:                                                # the loop body of extract_face, demo_model.py:22-29
            bb = b[1:] if len(b) == 5 else b
            _ = [int(p.numpy()) for p in (bb[0], bb[1], bb[0] + bb[2], bb[1] + bb[3])]
        n_boxes += b_ts.shape[0]
        assert b_ts.shape[0] == int(g["counts"][i])
    assert n_boxes > 0


@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("B,H,W", [(2, 480, 480), (3, 96, 132), (1, 50, 76), (5, 480, 640)])
def test_mbv3_stem_tensor_core_vs_torch(u8, B, H, W):
    """fd_mbv3_stem (Conv2dSame 3x3 stride 2, 3 -> 16, + bias + Hardswish; im2col patch tile + tcgen05, TMA zero fill =
    the TF "SAME" padding, ragged tiles at the right / bottom edge) against torch fp32 on bf16-rounded operands."""
    require_cuda()
    ops = fd().ops
    g = torch.Generator().manual_seed(B + H + W)
    if u8:
        x = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)
        xf = x.float() / 255.0
    else:
        x = torch.rand(B, 3, H, W, generator=g)
        xf = x
    w = torch.randn(16, 3, 3, 3, generator=g) * 0.3
    bias = torch.randn(16, generator=g) * 0.2
    ref = torch.nn.functional.hardswish(bo._conv_same(xf.bfloat16().float(), w.bfloat16().float(), 2)
                                        + bias.bfloat16().float().view(1, -1, 1, 1))
    Ho, Wo = ref.shape[-2:]
    out = torch.full((B, Ho, Wo, 16), 9.0, dtype=torch.bfloat16, device="cuda")
    ops.mbv3_stem(x.cuda(), w.cuda(), bias.cuda(), bo._same_pad(H, 3, 2)[0], bo._same_pad(W, 3, 2)[0], out)
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    assert rel_err(got, ref) <= 5e-3
