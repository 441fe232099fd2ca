"""The oracle (oracle/*.py) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import yolo_oracle as yo
from oracle import backbone_oracle as bo
from tests.util import GOLDEN, load_golden, seeded_poolresnet_params


def test_grid_encode_bit_exact():
    g = load_golden("grid_encode.npz")
    pos = 0
    for c, S in enumerate(g["S"]):
        b = g["boxes"][g["offsets"][c]:g["offsets"][c + 1]]
        want = g["fm"][pos:pos + 5 * S * S].reshape(5, S, S); pos += 5 * S * S
        got = yo.grid_encode(b, int(S), 480, 480)
        assert got.tobytes() == want.astype(np.float32).tobytes(), f"case {c}"


def test_decode_nms_bit_exact():
    g = load_golden("decode_nms.npz")
    assert len(g["count"]) > 30
    for c in range(len(g["count"])):
        Smap, Sdec, pthr, ithr = g["meta"][c]
        Smap, Sdec = int(Smap), int(Sdec)
        x = g["x"][c, :5 * Smap * Smap].reshape(5, Smap, Smap)
        want = g["out"][c, :g["count"][c]]
        got = yo.reduce_bounding_boxes(x, pthr, ithr, (3, 480, 480), Sdec)
        assert got.shape == want.shape, f"case {c}: {got.shape} vs {want.shape}"
        assert got.tobytes() == want.tobytes(), f"case {c}"


def test_nms_indices_bit_exact():
    g = load_golden("nms.npz")
    for c in range(len(g["n"])):
        n = int(g["n"][c])
        keep = yo.nms(g["boxes"][c, :n], g["scores"][c, :n], float(g["thr"][c]))
        want = g["keep"][c]; want = want[want >= 0]
        assert np.array_equal(keep, want), f"case {c}"


def test_yolo_loss_value_and_grad():
    g = load_golden("yolo_loss.npz")
    for c, S in enumerate(g["S"]):
        n = 5 * S * S
        p = g["pred"][c, :n].reshape(5, S, S); gt = g["gt"][c, :n].reshape(5, S, S)
        loss, d = yo.yolo_loss(p, gt)
        # the reference evaluates in f32; the oracle in f64: tolerance = f32 rounding of a ~S^2-term sum
        assert abs(loss - g["loss"][c]) <= 2e-6 * abs(g["loss"][c]) + 1e-6, f"case {c}"
        want = g["dpred"][c, :n].reshape(5, S, S)
        np.testing.assert_allclose(d, want, rtol=2e-5, atol=2e-6)
        # f32 evaluation of the same formula
        loss32, d32 = yo.yolo_loss(p, gt, dtype=np.float32)
        assert abs(loss32 - g["loss"][c]) <= 1e-5 * abs(g["loss"][c]) + 1e-6


def test_backbone_oracle_matches_reference_seeded():
    g = load_golden("backbone_seed2.npz")
    p = seeded_poolresnet_params(64, seed=2)
    for k, v in p.items():      # same weights as the reference module built under the same seed
        s = g["w_sum." + k]
        assert abs(v.double().sum().item() - s[0]) < 1e-9 and abs(v.double().abs().sum().item() - s[1]) < 1e-9, k
    x = torch.rand(2, 3, 480, 480, generator=torch.Generator().manual_seed(0))
    y = torch.from_numpy(g["y"])
    y_hat, loss, grads = bo.train_step(x, y, p, 10)
    assert torch.equal(y_hat, torch.from_numpy(g["y_hat"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    for k in p:
        assert abs(grads[k].double().norm().item() - float(g["g_norm." + k])) <= 1e-5 * float(g["g_norm." + k]) + 1e-12, k
        np.testing.assert_allclose(grads[k].reshape(-1)[:64].numpy(), g["g_head." + k], rtol=1e-4, atol=1e-7)


def test_backbone_oracle_official_checkpoint_demo_path():
    g = load_golden("official_medium.npz")
    p = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    for i in range(len(g["counts"])):
        x = torch.from_numpy(g["images"][i:i + 1]).float() / 255.0       # PoolResnet.py:94-97
        x = torch.cat([x, x])                                          # demo_model.py:20 stacks the frame twice
        head = bo.poolresnet_forward(x, p, 10)
        assert torch.equal(head[0], torch.from_numpy(g["heads"][i]))
        boxes = yo.reduce_bounding_boxes(head[0].numpy(), float(g["p_thr"]), float(g["iou_thr"]), (3, 480, 480), 10)
        want = g["boxes"][i, :g["counts"][i]]
        assert boxes.tobytes() == want.tobytes()


def test_resnet_oracle_matches_reference_seeded():
    """oracle resnet_forward (models/Resnet.py:89-99) against the real reference module: identical logits,
    loss and gradient fingerprints (tests/golden/make_golden_extra.py)."""
    g = load_golden("resnet_seed3.npz")
    p = seeded_poolresnet_params(64, seed=3, stem_k=3, stem_s=2, head_k=3)
    for k, v in p.items():
        s = g["w_sum." + k]
        assert abs(v.double().sum().item() - s[0]) < 1e-9 and abs(v.double().abs().sum().item() - s[1]) < 1e-9, k
    x = torch.rand(1, 3, 480, 480, generator=torch.Generator().manual_seed(4))
    y = torch.from_numpy(g["y"])
    y_hat, loss, grads = bo.train_step(x, y, p, 15, forward=bo.resnet_forward)
    assert torch.equal(y_hat, torch.from_numpy(g["y_hat"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    for k in p:
        assert abs(grads[k].double().norm().item() - float(g["g_norm." + k])) <= 1e-5 * float(g["g_norm." + k]) + 1e-12, k
        np.testing.assert_allclose(grads[k].reshape(-1)[:64].numpy(), g["g_head." + k], rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------ SSD rows (tests/golden/make_golden_extra.py)
def test_ssd_grid_encode_bit_exact():
    from oracle import ssd_oracle as so
    g = load_golden("ssd_encode.npz")
    for c in range(g["boxes"].shape[0]):
        b = g["boxes"][c, :g["counts"][c]]
        assert so.ssd_grid_encode(b, 480, 480).tobytes() == g["fm"][c].tobytes(), c


def test_ssd_decode_nms_bit_exact():
    from oracle import ssd_oracle as so
    g = load_golden("ssd_decode.npz")
    for c in range(g["x"].shape[0]):
        pthr, ithr, wp = g["cfg"][c]
        got = so.reduce_ssd_bounding_boxes(g["x"][c], pthr, ithr, (3, 480, 480), with_priors=bool(wp))
        want = g["out"][c, :g["counts"][c]]
        assert got.shape == want.shape and got.tobytes() == want.tobytes(), c


def test_ssd_loss_value_grad_and_mining_mask():
    from oracle import ssd_oracle as so
    g = load_golden("ssd_loss.npz")
    loss, dconf, dloc, mask = so.ssd_loss(g["conf"], g["loc"], g["labels"], g["gt_loc"], 10)
    assert np.array_equal(mask, g["mask"])
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    np.testing.assert_allclose(dconf, g["dconf"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(dloc, g["dloc"], rtol=2e-5, atol=1e-9)


def test_step_metrics_oracle_vs_torchvision_box_iou():
    """oracle.step_metrics restates ModelMeta.py:201-214 (third-party torchvision.ops.box_iou): pinned against the
    installed CPU op on integer-valued decoded rows, incl. zero-area boxes (0/0 -> nan -> 0) and empty sides."""
    import torch
    from torchvision.ops import box_iou
    from oracle import yolo_oracle as yo
    rng = np.random.default_rng(5)
    for case in range(20):
        ng, npred = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        def rows(n):
            r = np.zeros((n, 5), np.float32)
            r[:, 0] = rng.random(n)
            r[:, 1:3] = rng.integers(0, 400, (n, 2))
            r[:, 3:5] = rng.integers(0, 120, (n, 2))        # zero widths/heights occur
            return r
        g, p = rows(ng), rows(npred)
        if case % 5 == 0 and ng and npred:
            p[0] = g[0]; g[0, 3:5] = 0; p[0, 3:5] = 0        # identical zero-area pair: 0/0
        hits, s = yo.step_metrics(g, p, 0.5)
        if ng == 0 or npred == 0:
            assert hits == 0 and s == 0
            continue
        gt, pt = torch.from_numpy(g[:, 1:].copy()), torch.from_numpy(p[:, 1:].copy())
        gt[:, 2] += gt[:, 0]; gt[:, 3] += gt[:, 1]; pt[:, 2] += pt[:, 0]; pt[:, 3] += pt[:, 1]
        iou = torch.nan_to_num(box_iou(gt, pt), 0)
        assert hits == torch.where(iou > 0.5)[0].shape[0]
        assert abs(s - float(iou.double().sum())) <= 1e-6 * max(1.0, s)


def test_separable_oracle_matches_reference_seeded():
    """oracle separable_forward (models/SeparableCNN.py:40-51,104-117) against the REAL reference module
    (tests/golden/make_golden_sep.py): same seeded construction (weight fingerprints), identical head, and the
    decode + NMS rows through the hard-wired num_of_patches=16 on the 10x10 head (SeparableCNN.py:71) bit-exact."""
    import torch
    from oracle import backbone_oracle as bo
    from oracle import yolo_oracle as yo
    from tests.util import seeded_separable_params
    g = load_golden("separable_seed6.npz")
    p = seeded_separable_params(64, seed=6)
    for k, v in p.items():
        s = g["w_sum." + k]
        assert abs(v.double().sum().item() - s[0]) < 1e-9 and abs(v.double().abs().sum().item() - s[1]) < 1e-9, k
    x = torch.rand(2, 3, 480, 480, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        y = bo.separable_forward(x, p)
    assert tuple(y.shape) == (2, 5, 10, 10)
    assert (y - torch.from_numpy(g["y"])).abs().max().item() <= 1e-6
    for i in range(2):
        rows = yo.reduce_bounding_boxes(g["y"][i], 0.47, 0.3, (3, 480, 480), 16)
        k = int(g["counts"][i])
        assert rows.shape[0] == k and rows.tobytes() == g["boxes"][i, :k].tobytes()


def test_wide_models_oracle_matches_reference_seeded():
    """The 128-channel models (train_model.py:27-32 PoolResnet(filters=128); SeparableCNN.py:124 SeparableCNN(128))
    against the REAL reference (tests/golden/make_golden_wide.py): same seeded construction, identical head, loss and
    gradient norms -- pins the oracle the channel-plane engines are tested against."""
    import torch
    from oracle import backbone_oracle as bo
    from oracle import yolo_oracle as yo
    from tests.util import seeded_poolresnet_params, seeded_separable_params, synth_boxes
    g = load_golden("wide_seed12.npz")
    p = seeded_poolresnet_params(128, seed=12)
    for k, v in p.items():
        s = g["pool_w_sum." + k]
        assert abs(v.double().sum().item() - s[0]) < 1e-9 and abs(v.double().abs().sum().item() - s[1]) < 1e-9, k
    x = torch.rand(1, 3, 480, 480, generator=torch.Generator().manual_seed(13))
    boxes = synth_boxes(torch.Generator().manual_seed(14), 1, 100)
    y = torch.from_numpy(yo.grid_encode(boxes.numpy(), 10, 480, 480))[None]
    assert y.numpy().tobytes() == g["pool_y"].tobytes()
    y_hat, loss, grads = bo.train_step(x, y, p, 10)
    assert (y_hat - torch.from_numpy(g["pool_y_hat"])).abs().max().item() <= 1e-6
    assert abs(loss.item() - float(g["pool_loss"])) <= 1e-5 * abs(float(g["pool_loss"]))
    for k, v in grads.items():
        n = float(g["pool_g_norm." + k])
        assert abs(v.double().norm().item() - n) <= 1e-4 * max(n, 1e-12), k
    ps = seeded_separable_params(128, seed=11)
    for k, v in ps.items():
        s = g["sep_w_sum." + k]
        assert abs(v.double().sum().item() - s[0]) < 1e-9 and abs(v.double().abs().sum().item() - s[1]) < 1e-9, k
    xs = torch.rand(1, 3, 480, 480, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ys = bo.separable_forward(xs, ps)
    assert (ys - torch.from_numpy(g["sep_y_hat"])).abs().max().item() <= 1e-6


def test_mobilenetv3_oracle_matches_archive_golden():
    """oracle.backbone_oracle.mobilenetv3_forward (restated from the code stored in the official TorchScript archive, timm
    absent) == the archive's own heads on three of the 24 frames (make_golden_official.py): max-abs-diff <= 1e-6."""
    import cv2
    from oracle import backbone_oracle as bo
    g = np.load(os.path.join(GOLDEN, "official_mobilenetv3.npz"))
    im = np.load(os.path.join(GOLDEN, "official_images.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    for i in (0, 11, 23):
        rgb = cv2.imdecode(im["png"][im["offsets"][i]:im["offsets"][i + 1]], cv2.IMREAD_UNCHANGED)
        t = torch.from_numpy(rgb).permute(2, 0, 1)
        x = torch.stack([t, t]).float() / 255.0
        with torch.no_grad():
            y = bo.mobilenetv3_forward(x, sd)
        assert (y[0] - torch.from_numpy(g["heads"][i])).abs().max().item() <= 1e-6
        got = yo.reduce_bounding_boxes(g["heads"][i], float(g["p_thr"]), float(g["iou_thr"]), (3, 480, 480), 15)
        want = g["boxes"][i, :g["counts"][i]]
        assert got.shape == want.shape and got.tobytes() == want.tobytes()


def test_official_checkpoint_goldens_decode_with_oracle():
    """All 24 frames x (PoolResnet medium / small, Resnet medium): the oracle decode + NMS of the stored reference heads
    reproduces the stored demo-path boxes bit for bit; the backbone oracle reproduces two heads per model exactly."""
    import cv2
    from oracle import backbone_oracle as bo
    im = np.load(os.path.join(GOLDEN, "official_images.npz"))
    med = np.load(os.path.join(GOLDEN, "official_medium.npz"))
    for name, S, fwd in (("official_poolresnet_medium.npz", 10, bo.poolresnet_forward),
                         ("official_poolresnet_small.npz", 10, bo.poolresnet_forward),
                         ("official_resnet_medium.npz", 15, bo.resnet_forward)):
        g = np.load(os.path.join(GOLDEN, name))
        for i in range(24):
            got = yo.reduce_bounding_boxes(g["heads"][i], float(g["p_thr"]), float(g["iou_thr"]), (3, 480, 480), S)
            want = g["boxes"][i, :g["counts"][i]]
            assert got.shape == want.shape and got.tobytes() == want.tobytes(), (name, i)
        src = med if name == "official_poolresnet_medium.npz" else g
        sd = {k[3:]: torch.from_numpy(src[k]) for k in src.files if k.startswith("sd.")}
        for i in (3, 17):
            rgb = cv2.imdecode(im["png"][im["offsets"][i]:im["offsets"][i + 1]], cv2.IMREAD_UNCHANGED)
            t = torch.from_numpy(rgb).permute(2, 0, 1)
            x = torch.stack([t, t]).float() / 255.0
            with torch.no_grad():
                y = fwd(x, sd, S)
            assert (y[0] - torch.from_numpy(g["heads"][i])).abs().max().item() <= 1e-6, (name, i)


def test_ssd_model_oracle_matches_reference_golden():
    """oracle ssd_forward / ssd_train_step == the real reference's SSD model, loss and gradient norms (make_golden_ssd_model.py)."""
    from oracle import backbone_oracle as bo
    from tests.gpu_util import fd
    g = np.load(os.path.join(GOLDEN, "ssd_model_seed2.npz"))
    torch.manual_seed(2)
    m = fd().models.SSD.SSD(filters=16, input_shape=(3, 480, 480))          # the mirror's construction order == the reference's
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in p.items():
        assert abs(v.double().sum().item() - g["w_sum." + k][0]) < 1e-9, k
    x = torch.rand(2, 3, 480, 480, generator=torch.Generator().manual_seed(0))
    y_hat, loss, grads = bo.ssd_train_step(x[:1], torch.from_numpy(g["y"][:1]), p)      # one image keeps the CPU suite fast
    assert (y_hat - torch.from_numpy(g["y_hat"][:1])).abs().max().item() <= 1e-6
    y_hat2 = bo.ssd_forward(x[1:], p)
    assert (y_hat2 - torch.from_numpy(g["y_hat"][1:])).abs().max().item() <= 1e-6
