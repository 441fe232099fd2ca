"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference is pure Python; it is imported with five stub modules standing in
for packages that are not installed (albumentations, torchinfo, ptflops,
pytorch_lightning, gdown) -- none of them is on the hot path.  Everything that is
stored is an input/output pair of reference functions:

  decode_nms.npz   datasets/utils.py:95-170  ReduceBoundingBoxes.forward (+ torchvision nms)
  nms.npz          torchvision.ops.nms direct known-answer cases (ties, zero area, IoU == thr)
  grid_encode.npz  datasets/WIDERFace/dataset.py:32-64  convert_bbx_to_feature_map
  yolo_loss.npz    losses/YoloLoss.py:4-44  yolo_loss value + autograd gradient
  backbone_seed2.npz  models/PoolResnet.py  PoolResnet(filters=64,S=10) seeded weights:
                   logits, summed loss, gradient fingerprints for a B=2 batch
  official_medium.npz saved_models/official/PoolResnet/medium_model_10x10_480.pth weights
                   + demo-path outputs (demo_model.py:16-37) for two of imgs/test_imgs
"""
import io
import os
import sys
import types
import zipfile

import numpy as np
import torch
import torchvision  # noqa: F401  (registers torchvision::nms for jit.load)

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _LM(torch.nn.Module):
        def log(self, *a, **k):
            pass

    mod("albumentations")
    mod("albumentations.pytorch")
    mod("albumentations.pytorch.transforms", ToTensorV2=object)
    mod("torchinfo", summary=lambda *a, **k: None)
    mod("ptflops", get_model_complexity_info=lambda *a, **k: (0, 0))
    mod("pytorch_lightning", LightningModule=_LM, Trainer=object, LightningDataModule=object)
    mod("gdown")
    sys.path.insert(0, REF)


def synth_boxes(gen, kmin, kmax, size=480):
    """SURVEY 8d synthetic WIDERFace-like boxes: integer-valued (1,x,y,w,h) f32."""
    k = int(torch.randint(kmin, kmax + 1, (1,), generator=gen))
    x = torch.randint(0, size, (k,), generator=gen).float()
    y = torch.randint(0, size, (k,), generator=gen).float()
    lw = torch.rand(k, generator=gen) * (np.log(240.0) - np.log(4.0)) + np.log(4.0)
    lh = torch.rand(k, generator=gen) * (np.log(240.0) - np.log(4.0)) + np.log(4.0)
    w = torch.minimum(torch.round(torch.exp(lw)), size - x).clamp(min=1)
    h = torch.minimum(torch.round(torch.exp(lh)), size - y).clamp(min=1)
    return torch.stack([torch.ones(k), x, y, w, h], dim=1)


def main():
    install_stubs()
    from datasets.utils import ReduceBoundingBoxes
    from datasets.WIDERFace.dataset import WIDERFaceDataset
    from losses.YoloLoss import yolo_loss
    from models.PoolResnet import PoolResnet
    from torchvision.ops import nms

    gen = torch.Generator().manual_seed(1234)

    # ---------------- grid encode -------------------------------------------------------
    enc_boxes, enc_off, enc_fm, enc_S = [], [0], [], []
    for S, (kmin, kmax) in [(10, (1, 100)), (10, (1, 100)), (15, (101, 400)), (15, (1, 30)),
                            (10, (1, 1)), (16, (1, 60))]:
        ds = WIDERFaceDataset(None, S, (3, 480, 480))
        b = synth_boxes(gen, kmin, kmax)
        if S == 16:                      # boxes on/over the border: exercises the clamp
            b[0, 1:3] = torch.tensor([479.0, 479.0])
            b[1, 1:3] = torch.tensor([480.0, 0.0])
        fm = ds.convert_bbx_to_feature_map(b, (480, 480))
        enc_boxes.append(b.numpy()); enc_off.append(enc_off[-1] + b.shape[0])
        enc_fm.append(fm.numpy().reshape(-1)); enc_S.append(S)
    np.savez_compressed(os.path.join(OUT, "grid_encode.npz"),
                        boxes=np.concatenate(enc_boxes), offsets=np.array(enc_off),
                        S=np.array(enc_S), fm=np.concatenate(enc_fm))

    # ---------------- decode + NMS ------------------------------------------------------
    cases_x, cases_meta, cases_out, cases_cnt = [], [], [], []

    def add_case(x, pthr, ithr, S_dec):
        rb = ReduceBoundingBoxes(pthr, ithr, (3, 480, 480), S_dec)
        out = rb(x.clone()).numpy().astype(np.float32).reshape(-1, 5)
        cases_x.append(x.numpy().astype(np.float32).reshape(-1))
        cases_meta.append([x.shape[1], S_dec, pthr, ithr])
        cases_cnt.append(out.shape[0])
        pad = np.zeros((256, 5), np.float32); pad[:out.shape[0]] = out
        cases_out.append(pad)

    for S in (10, 15):
        for pthr, ithr in [(0.5, 0.5), (0.7, 0.01), (0.3, 0.3), (0.05, 0.9)]:
            for rep in range(3):
                x = torch.sigmoid(torch.randn(5, S, S, generator=gen) * 2.0)
                add_case(x, pthr, ithr, S)
    # nothing above threshold / everything above threshold
    add_case(torch.full((5, 10, 10), 0.1), 0.5, 0.5, 10)
    add_case(torch.full((5, 10, 10), 0.9), 0.5, 0.5, 10)
    # ties in score, zero-size boxes
    x = torch.sigmoid(torch.randn(5, 10, 10, generator=gen)); x[0] = 0.75; x[3:, :5] = 0.0
    add_case(x, 0.5, 0.5, 10)
    # threshold equality: conf == thr must NOT pass (strict >)
    x = torch.sigmoid(torch.randn(5, 10, 10, generator=gen)); x[0, ::2] = 0.5
    add_case(x, 0.5, 0.5, 10)
    # the SeparableCNN quirk: 10x10 map decoded with num_of_patches=16
    add_case(torch.sigmoid(torch.randn(5, 10, 10, generator=gen) * 2.0), 0.5, 0.5, 16)
    # encode -> decode round trip (dataset.py:125-139)
    for S in (10, 15):
        ds = WIDERFaceDataset(None, S, (3, 480, 480))
        fm = ds.convert_bbx_to_feature_map(synth_boxes(gen, 30, 60), (480, 480))
        add_case(fm, 0.5, 0.5, S)
        add_case(fm, 0.5, 1.0, S)
    maxlen = max(len(c) for c in cases_x)
    xs = np.zeros((len(cases_x), maxlen), np.float32)
    for i, c in enumerate(cases_x):
        xs[i, :len(c)] = c
    np.savez_compressed(os.path.join(OUT, "decode_nms.npz"), x=xs,
                        meta=np.array(cases_meta, np.float64), out=np.stack(cases_out),
                        count=np.array(cases_cnt))

    # ---------------- direct NMS known answers -------------------------------------------
    nb, ns, nt, nk, nn_ = [], [], [], [], []
    def add_nms(b, s, t):
        k = nms(b, s, t).numpy()
        pb = np.zeros((512, 4), np.float32); pb[:len(b)] = b.numpy()
        ps = np.zeros((512,), np.float32); ps[:len(s)] = s.numpy()
        pk = -np.ones((512,), np.int64); pk[:len(k)] = k
        nb.append(pb); ns.append(ps); nt.append(t); nk.append(pk); nn_.append(len(b))
    for n in (1, 2, 17, 100, 225, 400):
        for t in (0.5, 0.01, 0.3):
            xy = torch.randint(0, 400, (n, 2), generator=gen).float()
            wh = torch.randint(0, 120, (n, 2), generator=gen).float()     # zero sizes included
            s = torch.rand(n, generator=gen)
            s[::3] = s[0]                                                 # score ties
            add_nms(torch.cat([xy, xy + wh], 1), s, t)
    # IoU exactly equal to the threshold -> kept (strict >)
    add_nms(torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 5.]]), torch.tensor([0.9, 0.8]), 0.5)
    add_nms(torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 6.]]), torch.tensor([0.9, 0.8]), 0.5)
    # two zero-area boxes at the same spot -> 0/0 NaN -> both kept
    add_nms(torch.tensor([[5., 5., 5., 5.], [5., 5., 5., 5.]]), torch.tensor([0.9, 0.8]), 0.01)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), boxes=np.stack(nb), scores=np.stack(ns),
                        thr=np.array(nt), keep=np.stack(nk), n=np.array(nn_))

    # ---------------- yolo loss ------------------------------------------------------------
    lp, lg, ll, ld, lS = [], [], [], [], []
    for S in (10, 15, 10, 15, 10):
        ds = WIDERFaceDataset(None, S, (3, 480, 480))
        g = ds.convert_bbx_to_feature_map(synth_boxes(gen, 1, 100), (480, 480))
        p = torch.sigmoid(torch.randn(5, S, S, generator=gen)).requires_grad_(True)
        loss = yolo_loss(p, g)
        loss.backward()
        pad = lambda a: np.pad(a.reshape(-1), (0, 5 * 15 * 15 - a.size))
        lp.append(pad(p.detach().numpy())); lg.append(pad(g.numpy()))
        ll.append(loss.item()); ld.append(pad(p.grad.numpy())); lS.append(S)
    # NaN in pred (YoloLoss.py:8-9): replaced by 0.1, zero gradient
    S = 10
    ds = WIDERFaceDataset(None, S, (3, 480, 480))
    g = ds.convert_bbx_to_feature_map(synth_boxes(gen, 20, 40), (480, 480))
    p0 = torch.sigmoid(torch.randn(5, S, S, generator=gen)); p0[2, 3, 4] = float("nan"); p0[0, 0, 0] = float("nan")
    p = p0.clone().requires_grad_(True)
    loss = yolo_loss(p, g); loss.backward()
    pad = lambda a: np.pad(a.reshape(-1), (0, 5 * 15 * 15 - a.size))
    lp.append(pad(p.detach().numpy())); lg.append(pad(g.numpy())); ll.append(loss.item())
    ld.append(pad(p.grad.numpy())); lS.append(S)
    np.savez_compressed(os.path.join(OUT, "yolo_loss.npz"), pred=np.stack(lp), gt=np.stack(lg),
                        loss=np.array(ll, np.float64), dpred=np.stack(ld), S=np.array(lS))

    # ---------------- backbone, seeded weights ----------------------------------------------
    torch.manual_seed(2)
    model = PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10).eval()
    gx = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 480, 480, generator=gx)
    gb = torch.Generator().manual_seed(1)
    ds = WIDERFaceDataset(None, 10, (3, 480, 480))
    y = torch.stack([ds.convert_bbx_to_feature_map(synth_boxes(gb, 1, 100), (480, 480)) for _ in range(2)])
    y_hat = model(x)
    loss = 0
    for i in range(2):
        loss = loss + yolo_loss(y_hat[i], y[i])
    loss.backward()
    fp = {}
    for k, v in model.named_parameters():
        fp["w_sum." + k] = np.array([v.detach().double().sum().item(), v.detach().double().abs().sum().item()])
        g = v.grad.detach()
        fp["g_norm." + k] = np.array(g.double().norm().item())
        fp["g_head." + k] = g.reshape(-1)[:64].numpy().copy()
    for k in ("out.weight", "out.bias", "conv1.bias", "residual_blocks.0.conv1.bias",
              "residual_blocks.9.conv2.bias", "residual_blocks.5.conv1.weight"):
        fp["g_full." + k] = dict(model.named_parameters())[k].grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "backbone_seed2.npz"), y=y.numpy(), y_hat=y_hat.detach().numpy(),
                        loss=np.array(loss.item(), np.float64), **fp)

    # ---------------- official "medium" checkpoint, demo path --------------------------------
    src = os.path.join(REF, "saved_models/official/PoolResnet/medium_model_10x10_480.pth")
    buf = io.BytesIO()
    with zipfile.ZipFile(src) as zin, zipfile.ZipFile(buf, "w") as zout:
        for item in zin.infolist():
            data = zin.read(item.filename)
            if item.filename.endswith("functional_tensor.py"):
                t = data.decode()
                import re
                t = re.sub(r"ops\.torchvision\._interpolate_bi(linear|cubic)2d_aa\(img0, \[new_h, new_w\], False\)",
                           "img0", t)
                data = t.encode()
            zout.writestr(item, data)
    buf.seek(0)
    ts = torch.jit.load(buf, map_location="cpu").eval()
    sd = {k: v.detach().clone() for k, v in ts.state_dict().items()}
    eager = PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10,
                       probability_threshold=0.7, iou_threshold=0.01).eval()
    print("load_state_dict:", eager.load_state_dict(sd, strict=True))
    import cv2
    imgs, heads, boxes, counts = [], [], [], []
    for name in ("1.jpg", "12.jpg"):
        frame = cv2.imread(os.path.join(REF, "imgs/test_imgs", name))
        frame = cv2.resize(frame, (480, 480))                       # demo_model.py:18
        t = torch.tensor(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)).permute(2, 0, 1)   # :19
        t2 = torch.stack([t, t])                                    # :20
        with torch.no_grad():
            b_ts = ts(t2, predict=torch.tensor(1))                  # :21 (TorchScript archive)
            head = eager(t2.float() / 255.0)                        # raw head, same batch of 2 as the demo
            b_eager = eager(t2, predict=torch.tensor(1))
        assert torch.equal(b_ts, b_eager), "eager != TorchScript"
        imgs.append(t.numpy()); heads.append(head[0].numpy())
        pad = np.zeros((100, 5), np.float32); pad[:b_ts.shape[0]] = b_ts.numpy()
        boxes.append(pad); counts.append(b_ts.shape[0])
        print(name, "boxes:", b_ts.shape[0])
    np.savez_compressed(os.path.join(OUT, "official_medium.npz"),
                        images=np.stack(imgs), heads=np.stack(heads), boxes=np.stack(boxes),
                        counts=np.array(counts), p_thr=0.7, iou_thr=0.01,
                        **{"sd." + k: v.numpy() for k, v in sd.items()})
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
