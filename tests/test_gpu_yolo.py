"""GPU parity of the grid-head kernels (decode + NMS, YoloLoss, grid encode) through the C ABI,
against the reference-generated golden fixtures and against the oracle on seeded inputs.
Integer / index / rounded-coordinate work must be BIT-EXACT; the loss is checked to 1e-5 relative
(fp32 summation order differs from ATen's)."""
import numpy as np
import pytest
import torch

from oracle import yolo_oracle as yo
from tests.gpu_util import fd, require_cuda
from tests.util import load_golden, synth_boxes

pytestmark = pytest.mark.gpu


def _decode(x, pthr, ithr, Sdec):
    ops = fd().ops
    x = torch.as_tensor(x).cuda().float().contiguous()
    B, _, S1, S2 = x.shape
    boxes = torch.zeros((B, S1 * S2, 5), device="cuda")
    cells = torch.zeros((B, S1 * S2), dtype=torch.int32, device="cuda")
    counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    ops.decode_nms(x, pthr, ithr, 480, 480, Sdec, boxes, cells, counts)
    torch.cuda.synchronize()
    return boxes.cpu().numpy(), cells.cpu().numpy(), counts.cpu().numpy()


def test_decode_nms_golden_bit_exact():
    require_cuda()
    g = load_golden("decode_nms.npz")
    for c in range(len(g["count"])):
        Smap, Sdec, pthr, ithr = g["meta"][c]
        Smap, Sdec = int(Smap), int(Sdec)
        x = g["x"][c, :5 * Smap * Smap].reshape(1, 5, Smap, Smap)
        boxes, _, counts = _decode(x, pthr, ithr, Sdec)
        k = int(g["count"][c])
        assert counts[0] == k, f"case {c}: {counts[0]} vs {k}"
        assert boxes[0, :k].tobytes() == g["out"][c, :k].tobytes(), f"case {c}"


def test_decode_nms_reference_class_mirror():
    """Through the reference-shaped API: ReduceBoundingBoxes(...)(x[5,S,S]) and BaseModel.non_max_suppression."""
    require_cuda()
    RB = fd().datasets.utils.ReduceBoundingBoxes
    g = load_golden("decode_nms.npz")
    for c in (0, 5, 13, 24, 25):
        Smap, Sdec, pthr, ithr = g["meta"][c]
        Smap, Sdec = int(Smap), int(Sdec)
        x = torch.from_numpy(g["x"][c, :5 * Smap * Smap].reshape(5, Smap, Smap)).cuda()
        out = RB(pthr, ithr, (3, 480, 480), Sdec)(x)
        k = int(g["count"][c])
        assert tuple(out.shape) == (k, 5)
        assert out.cpu().numpy().tobytes() == g["out"][c, :k].tobytes()
    empty = RB(0.5, 0.5, (3, 480, 480), 10)(torch.zeros(5, 10, 10).cuda())
    assert tuple(empty.shape) == (0, 5)


@pytest.mark.parametrize("S,B", [(10, 64), (15, 256), (32, 8)])
def test_decode_nms_batch_vs_oracle(S, B):
    """Full batch sizes of BASELINE configs; oracle per image; kept cell indices bit-exact too."""
    require_cuda()
    gen = torch.Generator().manual_seed(7 + S)
    x = torch.sigmoid(torch.randn(B, 5, S, S, generator=gen) * 2.0)
    for pthr, ithr in [(0.5, 0.5), (0.7, 0.01)]:
        boxes, cells, counts = _decode(x, pthr, ithr, S)
        for b in range(0, B, max(1, B // 16)):
            want, wcell = yo.reduce_bounding_boxes(x[b].numpy(), pthr, ithr, (3, 480, 480), S, return_index=True)
            assert counts[b] == want.shape[0]
            assert boxes[b, :counts[b]].tobytes() == want.tobytes()
            assert np.array_equal(cells[b, :counts[b]], wcell)
        # size-independent properties on EVERY image: scores sorted descending, all above threshold,
        # idempotence of the kept set under a second NMS pass with the same threshold
        for b in range(B):
            k = counts[b]
            s = boxes[b, :k, 0]
            assert np.all(s[:-1] >= s[1:]) and np.all(s > np.float32(pthr))


def test_yolo_loss_golden():
    require_cuda()
    ops = fd().ops
    g = load_golden("yolo_loss.npz")
    for c, S in enumerate(g["S"]):
        n = 5 * S * S
        p = torch.from_numpy(g["pred"][c, :n].reshape(1, 5, S, S)).cuda()
        gt = torch.from_numpy(g["gt"][c, :n].reshape(1, 5, S, S)).cuda()
        loss = torch.zeros(1, device="cuda"); d = torch.zeros_like(p)
        ops.yolo_loss(p, gt, loss, None, d)
        assert abs(loss.item() - g["loss"][c]) <= 1e-5 * abs(g["loss"][c]) + 1e-6, f"case {c}"
        np.testing.assert_allclose(d.cpu().numpy().reshape(-1), g["dpred"][c, :n], rtol=2e-5, atol=2e-6)


def test_yolo_loss_autograd_mirror_and_batch():
    """losses.YoloLoss.yolo_loss(pred[5,S,S], gt) keeps the reference signature and autograd contract."""
    require_cuda()
    L = fd().losses.YoloLoss
    gen = torch.Generator().manual_seed(3)
    B, S = 64, 10
    gts = torch.stack([torch.from_numpy(yo.grid_encode(synth_boxes(gen, 1, 100).numpy(), S, 480, 480))
                       for _ in range(B)]).cuda()
    pred = torch.sigmoid(torch.randn(B, 5, S, S, generator=gen)).cuda().requires_grad_(True)
    total = L.yolo_loss_batch(pred, gts)
    total.backward()
    want_loss, want_d = 0.0, []
    for b in range(B):
        l, d = yo.yolo_loss(pred[b].detach().cpu().numpy(), gts[b].cpu().numpy())
        want_loss += l; want_d.append(d)
    assert abs(total.item() - want_loss) <= 1e-5 * want_loss
    np.testing.assert_allclose(pred.grad.cpu().numpy(), np.stack(want_d), rtol=3e-5, atol=3e-6)
    # per-image reference signature, scaled upstream gradient
    p1 = pred[3].detach().clone().requires_grad_(True)
    (2.5 * L.yolo_loss(p1, gts[3])).backward()
    np.testing.assert_allclose(p1.grad.cpu().numpy(), 2.5 * want_d[3], rtol=3e-5, atol=3e-6)


def test_grid_encode_golden_bit_exact():
    require_cuda()
    enc = fd().datasets.WIDERFace.dataset
    g = load_golden("grid_encode.npz")
    pos = 0
    for c, S in enumerate(g["S"]):
        b = torch.from_numpy(g["boxes"][g["offsets"][c]:g["offsets"][c + 1]])
        want = g["fm"][pos:pos + 5 * S * S].reshape(5, S, S); pos += 5 * S * S
        got = enc.WIDERFaceDataset(None, int(S), (3, 480, 480)).convert_bbx_to_feature_map(b.cuda(), (480, 480))
        assert got.cpu().numpy().tobytes() == want.astype(np.float32).tobytes(), f"case {c}"


def test_grid_encode_ragged_batch_and_roundtrip():
    """Ragged batch incl. an EMPTY image; encode -> decode (iou_thr 1.0) returns the boxes of the
    winning cells (the reference's own round-trip assertion, dataset.py:125-139)."""
    require_cuda()
    enc = fd().datasets.WIDERFace.dataset
    gen = torch.Generator().manual_seed(11)
    S = 15
    boxes = [synth_boxes(gen, 101, 400) for _ in range(6)] + [torch.zeros((0, 5))] + [synth_boxes(gen, 1, 3)]
    fm = enc.convert_bbx_to_feature_map_batch(boxes, S, (480, 480), device=torch.device("cuda"))
    for i, b in enumerate(boxes):
        want = yo.grid_encode(b.numpy(), S, 480, 480)
        assert fm[i].cpu().numpy().tobytes() == want.tobytes(), f"image {i}"
    out, _, counts = _decode(fm.cpu(), 0.5, 1.0, S)
    for i, b in enumerate(boxes):
        want = yo.grid_encode(b.numpy(), S, 480, 480)
        occupied = int((want[0] > 0.5).sum())
        assert counts[i] == occupied


def test_step_metrics_batch_vs_oracle_and_reference_loop():
    """fd_box_metrics (ModelMeta.py:184-214 batched): hit counts bit-exact against the oracle restatement of
    torchvision.box_iou, IoU sums to 1e-5; ModelMeta.step's batched metrics == the reference's per-image loop."""
    require_cuda()
    pkg = fd()
    ops, RB = pkg.ops, pkg.datasets.utils.ReduceBoundingBoxes
    gen = torch.Generator().manual_seed(21)
    B, S = 48, 10
    boxes = [synth_boxes(gen, 1, 100) for _ in range(B - 2)] + [torch.zeros((0, 5)), synth_boxes(gen, 1, 2)]
    y = pkg.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, S, (480, 480), device=torch.device("cuda"))
    noise = torch.rand(y.shape, generator=gen).cuda()
    y_hat = (y * (0.75 + 0.25 * noise) + (1 - y[:, :1]) * noise * 0.6).clamp(0, 1)   # jittered boxes + false positives
    y_hat[3] = 0.0                                                                  # an image without predictions
    rb = RB(0.5, 0.5, (3, 480, 480), S)
    gb, gn = rb.batch_forward(y)
    pb, pn = rb.batch_forward(y_hat)
    m = torch.empty((B, 4), device="cuda")
    ops.box_metrics(gb, gn, pb, pn, 0.5, m)
    m = m.cpu().numpy()
    gb, gn, pb, pn = gb.cpu().numpy(), gn.cpu().numpy(), pb.cpu().numpy(), pn.cpu().numpy()
    for i in range(B):
        hits, s = yo.step_metrics(gb[i, :gn[i]], pb[i, :pn[i]], 0.5)
        assert int(m[i, 0]) == hits, f"image {i}"
        assert abs(m[i, 1] - s) <= 1e-5 * max(1.0, abs(s)), f"image {i}"
        assert (int(m[i, 2]), int(m[i, 3])) == (gn[i], pn[i])
    assert pn[3] == 0 and m[3, 0] == 0

    class _Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.reduce_bounding_boxes = rb
            self.w = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x):
            return y_hat + 0 * self.w

        def non_max_suppression(self, x):
            bx, n = rb.batch_forward(x)
            return rb.batch_to_tuple(bx, n)

    meta = pkg.models.ModelMeta(_Holder().cuda())        # `from models import ModelMeta` as in train_model.py:6
    out = meta.step((None, y, None), 1)
    ti, tr, tp = meta._metrics_per_image(y, y_hat)
    n = B
    assert abs(out["total_recall"] - tr / n) <= 1e-9 and abs(out["total_precision"] - tp / n) <= 1e-9
    assert abs(float(out["total_iou"]) - float(ti) / n) <= 1e-4 * max(1.0, float(ti) / n)
