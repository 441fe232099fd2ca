"""Golden fixture of the reference's SSD MODEL (models/SSD.py) -- run in the build container only:

    python tests/golden/make_golden_ssd_model.py

ssd_model_seed2.npz: SSD(filters=16, (3,480,480)) under torch.manual_seed(2), eval mode, batch of 2 seeded images,
targets from the reference's own multi-scale encoder (datasets/WIDERFace/dataset_ssd.py:36-76,134-139):
weight fingerprints (construction order / seeding parity), y_hat [2,4774,5], ssd_loss(.., 10) value (ModelMetaSSD.py:175)
and the gradient norm of every parameter + a few full gradients."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    mg.install_stubs()
    from datasets.WIDERFace.dataset_ssd import WIDERFaceDatasetSSD
    from losses.SSDLoss import ssd_loss
    from models.SSD import SSD
    torch.manual_seed(2)
    model = SSD(filters=16, input_shape=(3, 480, 480)).eval()
    x = torch.rand(2, 3, 480, 480, generator=torch.Generator().manual_seed(0))
    gb = torch.Generator().manual_seed(1)
    ds = WIDERFaceDatasetSSD(None, 10, (3, 480, 480))
    ys, boxes = [], []
    for _ in range(2):
        b = mg.synth_boxes(gb, 5, 60)
        boxes.append(b.numpy())
        fms = [ds.convert_bbx_to_feature_map(b, (480, 480), ps).permute(1, 2, 0).reshape(ps * ps, 5)
               for ps in (60, 30, 15, 7)]                                   # dataset_ssd.py:134-139
        ys.append(torch.cat(fms, dim=0))
    y = torch.stack(ys)
    y_hat = model(x)
    loss = ssd_loss(y_hat[:, :, 0], y_hat[:, :, 1:], y[:, :, 0], y[:, :, 1:], 10)
    loss.backward()
    fp = {}
    for k, v in model.named_parameters():
        fp["w_sum." + k] = np.array([v.detach().double().sum().item(), v.detach().double().abs().sum().item()])
        fp["g_norm." + k] = np.array(v.grad.double().norm().item())
    for k in ("input_normalizer.weight", "feature_extractor.0.pointwise_conv_skip.weight", "feature_extractor.4.conv2.bias",
              "continue_layers.1.0.conv1.bias", "extracting_layers.2.0.weight", "extracting_layers.0.0.bias"):
        fp["g_full." + k] = dict(model.named_parameters())[k].grad.numpy().copy()
    kmax = max(b.shape[0] for b in boxes)
    bp = np.zeros((2, kmax, 5), np.float32)
    for i, b in enumerate(boxes):
        bp[i, :b.shape[0]] = b
    np.savez_compressed(os.path.join(HERE, "ssd_model_seed2.npz"), y=y.numpy(), y_hat=y_hat.detach().numpy(),
                        loss=np.array(loss.item(), np.float64), boxes=bp, box_counts=np.array([b.shape[0] for b in boxes]),
                        **fp)
    print("loss", loss.item(), "size", os.path.getsize(os.path.join(HERE, "ssd_model_seed2.npz")))


if __name__ == "__main__":
    main()
