"""GPU parity of the SSD head path (grid encode, decode + NMS, ssd_loss) through the C ABI / the reference-shaped
Python mirrors, against golden outputs of the REAL reference (tests/golden/make_golden_extra.py) and against the
numpy oracle at batch sizes the reference would need minutes for.  Integer / index work is bit-exact; the loss is
compared with a relative tolerance of 2e-5 (fp32 sums of ~7600 log terms in a different order)."""
import numpy as np
import pytest
import torch

from oracle import ssd_oracle as so
from tests.gpu_util import fd, require_cuda
from tests.util import load_golden, synth_boxes

pytestmark = pytest.mark.gpu


def test_ssd_grid_encode_golden_and_batch():
    require_cuda()
    g = load_golden("ssd_encode.npz")
    enc = fd().datasets.WIDERFace.dataset_ssd
    boxes = [torch.from_numpy(g["boxes"][c, :g["counts"][c]]) for c in range(g["boxes"].shape[0])]
    out = enc.convert_bbx_to_feature_maps_batch(boxes, (480, 480)).cpu().numpy()
    for c in range(len(boxes)):
        assert out[c].tobytes() == g["fm"][c].tobytes(), c
    gen = torch.Generator().manual_seed(31)
    many = [synth_boxes(gen, 1, 119) for _ in range(128)]          # config 5: batch 128, < 120 faces per image
    out = enc.convert_bbx_to_feature_maps_batch(many, (480, 480)).cpu().numpy()
    for c in (0, 17, 127):
        assert out[c].tobytes() == so.ssd_grid_encode(many[c].numpy(), 480, 480).tobytes()
    # the reference's own round trip (dataset_ssd.py:142-150): decode(encode(boxes)) returns the boxes
    red = fd().datasets.utils.ReduceSSDBoundingBoxes(0.5, 0.5, (3, 480, 480), with_priors=True)
    dec = red(torch.from_numpy(out[3]).cuda()).cpu()
    want = so.reduce_ssd_bounding_boxes(out[3], 0.5, 0.5, (3, 480, 480), with_priors=True)
    assert dec.numpy().tobytes() == want.tobytes()


def test_ssd_decode_nms_golden_bit_exact():
    require_cuda()
    g = load_golden("ssd_decode.npz")
    R = fd().datasets.utils.ReduceSSDBoundingBoxes
    for c in range(g["x"].shape[0]):
        pthr, ithr, wp = g["cfg"][c]
        red = R(float(pthr), float(ithr), (3, 480, 480), (60, 30, 15, 7), with_priors=bool(wp))
        got = red(torch.from_numpy(g["x"][c]).cuda()).cpu().numpy()
        want = g["out"][c, :g["counts"][c]]
        assert got.shape == want.shape and got.tobytes() == want.tobytes(), c
    # batched, empty and dense rows in one launch
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(8, 4774, 5, generator=gen)
    x[:, :, 0] = torch.sigmoid(torch.randn(8, 4774, generator=gen) * 2 - 3)
    x[:, :, 3:] *= 0.3
    x[1, :, 0] = 0.1                                   # nothing above the threshold
    x[2, :, 0] = torch.rand(4774, generator=gen)       # ~2400 candidates
    red = R(0.5, 0.5, (3, 480, 480), with_priors=True)
    boxes, counts = red.batch_forward(x.cuda())
    boxes, counts = boxes.cpu().numpy(), counts.cpu().numpy()
    for b in range(8):
        want = so.reduce_ssd_bounding_boxes(x[b].numpy(), 0.5, 0.5, (3, 480, 480), with_priors=True)
        assert counts[b] == want.shape[0], b
        assert boxes[b, :counts[b]].tobytes() == want.tobytes(), b
    assert counts[1] == 0


def test_ssd_loss_golden_and_autograd():
    require_cuda()
    g = load_golden("ssd_loss.npz")
    L = fd().losses.SSDLoss
    conf = torch.from_numpy(g["conf"]).cuda().requires_grad_(True)
    loc = torch.from_numpy(g["loc"]).cuda().requires_grad_(True)
    labels, gt_loc = torch.from_numpy(g["labels"]).cuda(), torch.from_numpy(g["gt_loc"]).cuda()
    loss = L.ssd_loss(conf, loc, labels, gt_loc, 10)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
    np.testing.assert_allclose(conf.grad.cpu().numpy(), g["dconf"], rtol=2e-4, atol=1e-8)
    np.testing.assert_allclose(loc.grad.cpu().numpy(), g["dloc"], rtol=2e-4, atol=1e-8)
    mask = L.hard_negative_mining(-torch.log(torch.from_numpy(g["conf"]).cuda()), labels, 10).cpu().numpy()
    assert np.array_equal(mask, g["mask"])             # the mined set, bit-exact
    # ties in the ranking (constant confidences): lower prior index wins, like the reference's stable sort
    B, P = 3, 4774
    conf_t = torch.full((B, P), 0.25)
    lab = torch.zeros(B, P); lab[:, [5, 100, 4000]] = 0.94
    want = so.hard_negative_mining(-np.log(conf_t.numpy()), lab.numpy(), 10)
    got = L.hard_negative_mining(-torch.log(conf_t).cuda(), lab.cuda(), 10).cpu().numpy()
    assert np.array_equal(got, want) and got.sum() == B * 33
    # batch 128 (config 5) against the oracle
    gen = torch.Generator().manual_seed(77)
    B = 128
    conf_b = torch.sigmoid(torch.randn(B, P, generator=gen) * 1.5 - 1.0)
    loc_b = torch.rand(B, P, 4, generator=gen) * 1.4 - 0.2
    enc = fd().datasets.WIDERFace.dataset_ssd
    gt = enc.convert_bbx_to_feature_maps_batch([synth_boxes(gen, 1, 119) for _ in range(B)], (480, 480))
    lo, dcf, dlc, msk = so.ssd_loss(conf_b.numpy(), loc_b.numpy(), gt[:, :, 0].cpu().numpy(),
                                    gt[:, :, 1:].cpu().numpy(), 10)
    cb = conf_b.cuda().requires_grad_(True); lb_ = loc_b.cuda().requires_grad_(True)
    out = L.ssd_loss(cb, lb_, gt[:, :, 0], gt[:, :, 1:], 10)
    out.backward()
    assert abs(out.item() - lo) <= 2e-5 * abs(lo)
    np.testing.assert_allclose(cb.grad.cpu().numpy(), dcf, rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(lb_.grad.cpu().numpy(), dlc, rtol=2e-4, atol=1e-9)
