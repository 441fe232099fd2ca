"""Golden fixtures of the reference's four official checkpoints over ALL 24 ``imgs/test_imgs`` (BASELINE config 1, the
``demo_model.py`` path).  Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_official.py

For every archive (patched for the dead ``_interpolate_bilinear2d_aa`` branch) and every image:
``cv2.resize(480,480)`` -> RGB -> uint8 [3,480,480] -> stacked twice (demo_model.py:18-20) ->
``model(t2, predict=torch.tensor(1))`` (demo_model.py:21) = the boxes of image 0, plus the raw sigmoid head of the same
batch.  The frames are stored ONCE, losslessly (PNG bytes, ``official_images.npz``), the weights per archive.

  official_images.npz             24 PNG-encoded 480x480 RGB frames + file names
  official_poolresnet_medium.npz  PoolResnet F=64 S=10 : heads [24,5,10,10], boxes, counts (weights: official_medium.npz)
  official_poolresnet_small.npz   PoolResnet F=32 S=10 : + state dict
  official_resnet_medium.npz      Resnet F=64 S=15     : + state dict
  official_mobilenetv3.npz        MobilenetV3 S=15     : + state dict (incl. BatchNorm statistics)
"""
import io
import os
import re
import sys
import zipfile

import cv2
import numpy as np
import torch
import torchvision  # noqa: F401  (registers torchvision::nms for jit.load)

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_archive(path):
    buf = io.BytesIO()
    with zipfile.ZipFile(path) as zin, zipfile.ZipFile(buf, "w") as zout:
        for item in zin.infolist():
            data = zin.read(item.filename)
            if item.filename.endswith("functional_tensor.py"):
                t = data.decode()
                t = re.sub(r"ops\.torchvision\._interpolate_bi(linear|cubic)2d_aa\(img0, \[new_h, new_w\], False\)",
                           "img0", t)
                data = t.encode()
            zout.writestr(item, data)
    buf.seek(0)
    return torch.jit.load(buf, map_location="cpu").eval()


def main():
    d = os.path.join(REF, "imgs/test_imgs")
    names = sorted(os.listdir(d), key=lambda s: int(s.split(".")[0]))
    frames, pngs = [], []
    for n in names:
        frame = cv2.resize(cv2.imread(os.path.join(d, n)), (480, 480))          # demo_model.py:18
        rgb = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)                             # :19
        ok, png = cv2.imencode(".png", rgb, [cv2.IMWRITE_PNG_COMPRESSION, 9])    # stored as RGB planes-last, lossless
        assert ok and np.array_equal(cv2.imdecode(png, cv2.IMREAD_UNCHANGED), rgb)
        frames.append(torch.tensor(rgb).permute(2, 0, 1))
        pngs.append(np.frombuffer(png.tobytes(), np.uint8))
    lens = np.array([len(p) for p in pngs])
    blob = np.concatenate(pngs)
    np.savez(os.path.join(OUT, "official_images.npz"), png=blob, offsets=np.concatenate([[0], np.cumsum(lens)]),
             names=np.array(names))

    jobs = [("official_poolresnet_medium.npz", "PoolResnet/medium_model_10x10_480.pth", False),
            ("official_poolresnet_small.npz", "PoolResnet/small_model_10x10_480.pth", True),
            ("official_resnet_medium.npz", "Resnet/medium_model_15x15_480.pth", True),
            ("official_mobilenetv3.npz", "MobilenetV3Backbone/medium_model_15x15_480.pth", True)]
    for out_name, rel, with_sd in jobs:
        ts = load_archive(os.path.join(REF, "saved_models/official", rel))
        rb = ts.reduce_bounding_boxes          # the thresholds BAKED into the archive (0.7 / 0.01), not the class defaults
        heads, boxes, counts = [], [], []
        for t in frames:
            t2 = torch.stack([t, t])                                             # demo_model.py:20
            with torch.no_grad():
                b = ts(t2, predict=torch.tensor(1))                              # :21
                h = ts(t2.float() / 255.0)                                       # raw head of the same batch
            heads.append(h[0].numpy())
            S2 = h.shape[2] * h.shape[3]
            pad = np.zeros((S2, 5), np.float32)
            pad[:b.shape[0]] = b.numpy()
            boxes.append(pad)
            counts.append(b.shape[0])
        extra = {}
        if with_sd:
            extra = {"sd." + k: v.numpy() for k, v in ts.state_dict().items()}
        np.savez_compressed(os.path.join(OUT, out_name), heads=np.stack(heads), boxes=np.stack(boxes),
                            counts=np.array(counts), p_thr=float(rb.probability_threshold),
                            iou_thr=float(rb.iou_threshold), **extra)
        print(out_name, "boxes per image:", counts, "thr", float(rb.probability_threshold), float(rb.iou_threshold))
    for f in sorted(os.listdir(OUT)):
        if f.startswith("official_"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
