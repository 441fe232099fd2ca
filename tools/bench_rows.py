"""Device timings of the SURVEY 8 rows that are not the headline bench line (one B200, CUDA events around CUDA-graph
replays, inputs resident in HBM).  Prints one JSON object per row; `python tools/bench_rows.py > profiles/...`.

  row 3   Resnet (standard) train step, dense-crowd targets (BASELINE config 3, per-GPU batch 16)
  row 9   SeparableCNN(filters=64) inference + decode + NMS, batch 256 (BASELINE config 4)
  rows 4,5,7  batched decode + threshold + NMS on a synthetic sigmoid(N(0,2)) head, batch 256, S=15
  row 6   YoloLoss value + gradient, batch 64      row 8   grid-cell assignment, batch 64, <= 100 boxes
  rows 12,13  ssd_loss (mining + BCE + smooth-L1 + grads) and SSD decode + NMS, batch 128, 4774 priors
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import synth_boxes  # noqa: E402

fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbps_sustained", 6436.0) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6436.0


def timed(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = fd.native.launch_count()
    with torch.cuda.graph(g):
        fn()
    launches = fd.native.launch_count() - n0
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, launches      # us, launches


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(1)

    # ---- row 3: Resnet standard, S=15, dense crowd
    B = 16
    torch.manual_seed(3)
    m = fd.models.Resnet.Resnet(filters=64, input_shape=(3, 480, 480), num_of_patches=15).cuda().train()
    eng = m.engine
    eng.bind(dict(m.named_parameters()))
    x = torch.rand(B, 3, 480, 480, generator=gen).cuda()
    gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch([synth_boxes(gen, 101, 400) for _ in range(B)], 15,
                                                                        (480, 480), device=dev)
    opt = fd.optim.FlatAdam(eng, lr=1e-4, capturable=True)
    us, n = timed(lambda: eng.train_step(x, gt, dropout=True, optimizer=opt), reps=10)
    flops = 3 * 11.68e9 * B
    emit(row=3, what="Resnet(filters=64, S=15) train step: forward + summed YoloLoss + backward + Adam, batch 16, "
                     "101..400 boxes/image, train-mode dropout", us=us, launches=n, images_per_s=B / us * 1e6,
         tflops=flops / us / 1e6)
    m.eval()
    red = m.reduce_bounding_boxes
    us, n = timed(lambda: red.batch_forward(eng.forward(x, train=False, dropout=False).y), reps=10)
    emit(row=3, what="Resnet(filters=64, S=15) inference + decode + NMS, batch 16", us=us, launches=n,
         images_per_s=B / us * 1e6, tflops=11.68e9 * B / us / 1e6)
    del m, eng, x, gt, opt
    torch.cuda.empty_cache()

    # ---- rows 1/2 at the width train_model.py trains: PoolResnet(filters=128) on two 64-channel planes (PlanarEngine)
    B = 64
    torch.manual_seed(4)
    m = fd.models.PoolResnet.PoolResnet(filters=128, input_shape=(3, 480, 480), num_of_patches=10).cuda().train()
    eng = m.engine
    eng.bind(dict(m.named_parameters()))
    x = torch.rand(B, 3, 480, 480, generator=gen).cuda()
    gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch([synth_boxes(gen, 1, 100) for _ in range(B)], 10,
                                                                        (480, 480), device=dev)
    topt = m.flat_optimizer(lr=1e-4, capturable=True)
    topt._ensure_state()
    us, n = timed(lambda: eng.train_step(x, gt, dropout=True, optimizer=topt), reps=10)
    emit(row="1,2", what="PoolResnet(filters=128, S=10) train step: forward + summed YoloLoss + backward + Adam (fd_adam_flat), batch 64, "
                         "train-mode dropout (PlanarEngine: two channel planes, cta_group::2 wide kernels for the 3x3 convolutions)", us=us, launches=n,
         images_per_s=B / us * 1e6, tflops=3 * 3.997e9 * B / us / 1e6)
    m.eval()
    red = m.reduce_bounding_boxes
    us, n = timed(lambda: red.batch_forward(eng.forward(x, train=False, dropout=False).y), reps=10)
    emit(row="1,2", what="PoolResnet(filters=128, S=10) inference + decode + NMS, batch 64", us=us, launches=n,
         images_per_s=B / us * 1e6, tflops=3.997e9 * B / us / 1e6)
    del m, eng, x, gt, topt
    torch.cuda.empty_cache()

    # ---- row 9: SeparableCNN
    B = 256
    torch.manual_seed(6)
    sm = fd.models.SeparableCNN.SeparableCNN(filters=64, input_shape=(3, 480, 480)).cuda().eval()
    sm.engine.bind(dict(sm.named_parameters()))
    x = torch.rand(B, 3, 480, 480, device=dev)
    red = sm.reduce_bounding_boxes
    us, n = timed(lambda: red.batch_forward(sm.engine.forward(x)), reps=10)
    emit(row=9, what="SeparableCNN(filters=64) inference + decode + NMS, batch 256, fp32 images resident in HBM", us=us,
         launches=n, images_per_s=B / us * 1e6,
         hbm_floor_us=(B * 3 * 480 * 480 * 4 + 2 * B * 128 * (3600 + 900 + 8 * 225)) / HBM / 1e3)
    del sm
    torch.manual_seed(7)
    sm = fd.models.SeparableCNN.SeparableCNN(filters=128, input_shape=(3, 480, 480)).cuda().eval()
    sm.engine.bind(dict(sm.named_parameters()))
    us, n = timed(lambda: red.batch_forward(sm.engine.forward(x)), reps=5)
    emit(row=9, what="SeparableCNN(filters=128) inference + decode + NMS, batch 256 (channel planes: pointwise convolutions on "
                     "fd_conv3x3_wide in centre-tap mode + fd_dwconv3x3_lrelu)", us=us, launches=n, images_per_s=B / us * 1e6)
    w_pw = (torch.randn(2, 64, 64, device=dev) * 0.1).bfloat16()
    w_dw = torch.randn(9, 64, device=dev) * 0.3
    for H in (60, 30, 15):
        xs = [torch.randn(B, H, H, 64, device=dev).bfloat16() for _ in range(4 if H == 60 else 8)]
        outs = [torch.empty_like(t) for t in xs]

        def run():
            for t, o in zip(xs, outs):
                ops.sepblock_fwd(t, w_pw[0], w_dw, w_pw[1], 0.2, o)
        us, n = timed(run)
        us /= len(xs)
        byt = B * H * H * 256
        emit(row=9, what=f"fd_sepblock_fwd {H}x{H}, batch 256 (rotating buffers > L2)", us=us,
             algorithmic_bytes=byt, gbps=byt / us / 1e3, hbm_frac=byt / us / 1e3 / HBM)
    del sm, x, xs, outs
    torch.cuda.empty_cache()

    # ---- rows 4/5/7: decode + NMS
    B, S = 256, 15
    head = torch.sigmoid(torch.randn(B, 5, S, S, generator=gen) * 2).cuda()
    for pthr, ithr in ((0.5, 0.5), (0.7, 0.01)):
        rb = fd.datasets.utils.ReduceBoundingBoxes(pthr, ithr, (3, 480, 480), S)
        us, n = timed(lambda: rb.batch_forward(head))
        _, counts = rb.batch_forward(head)
        emit(row="4,5,7", what=f"fd_decode_nms batch 256, S=15, p_thr {pthr}, iou_thr {ithr}", us=us, launches=n,
             images_per_s=B / us * 1e6, kept=int(counts.sum().item()), algorithmic_bytes=B * 5 * S * S * 4)

    # ---- row 6: YoloLoss + gradient ; row 8: grid encode
    B, S = 64, 10
    boxes = [synth_boxes(gen, 1, 100) for _ in range(B)]
    enc = fd.datasets.WIDERFace.dataset
    gt = enc.convert_bbx_to_feature_map_batch(boxes, S, (480, 480), device=dev)
    pred = torch.sigmoid(torch.randn(B, 5, S, S, generator=gen)).cuda()
    loss, dpred = torch.empty(B, device=dev), torch.empty_like(pred)
    us, n = timed(lambda: ops.yolo_loss(pred, gt, loss, None, dpred))
    emit(row=6, what="fd_yolo_loss value + gradient, batch 64, S=10", us=us, launches=n, algorithmic_bytes=3 * B * 5 * S * S * 4)
    flat = torch.cat(boxes).cuda()
    offs = torch.tensor([0] + list(torch.tensor([b.shape[0] for b in boxes]).cumsum(0)), dtype=torch.int32).cuda()
    out = torch.empty((B, 5, S, S), device=dev)
    us, n = timed(lambda: ops.grid_encode(flat, offs, S, 480, 480, out))
    emit(row=8, what="fd_grid_encode batch 64, 1..100 boxes/image, S=10", us=us, launches=n, boxes=int(flat.shape[0]))

    # ---- rows 12/13: SSD
    B, P = 128, 4774
    conf = torch.sigmoid(torch.randn(B, P, generator=gen) * 1.5 - 1.0).cuda()
    loc = (torch.rand(B, P, 4, generator=gen) * 1.4 - 0.2).cuda()
    encs = fd.datasets.WIDERFace.dataset_ssd
    gts = encs.convert_bbx_to_feature_maps_batch([synth_boxes(gen, 1, 119) for _ in range(B)], (480, 480))
    labels, gloc = gts[:, :, 0].contiguous(), gts[:, :, 1:].contiguous()
    sums, npos = torch.empty(B, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
    dconf, dloc = torch.empty_like(conf), torch.empty_like(loc)
    us, n = timed(lambda: ops.ssd_loss(conf, loc, labels, gloc, 10, sums, npos, None, dconf, dloc))
    byt = B * P * 4 * (1 + 4 + 1 + 4 + 1 + 4)
    emit(row=12, what="fd_ssd_loss (hard-negative mining + BCE + smooth-L1 + gradients), batch 128 x 4774 priors", us=us,
         launches=n, algorithmic_bytes=byt, gbps=byt / us / 1e3, hbm_frac=byt / us / 1e3 / HBM)
    xs = torch.rand(B, P, 5, generator=gen)
    xs[:, :, 0] = torch.sigmoid(torch.randn(B, P, generator=gen) * 2 - 3)
    xs[:, :, 3:] *= 0.3
    xs = xs.cuda()
    red = fd.datasets.utils.ReduceSSDBoundingBoxes(0.5, 0.5, (3, 480, 480))
    us, n = timed(lambda: red.batch_forward(xs))
    _, counts = red.batch_forward(xs)
    emit(row=13, what="fd_ssd_decode_nms batch 128 x 4774 priors, p_thr 0.5, iou_thr 0.5", us=us, launches=n,
         images_per_s=B / us * 1e6, kept=int(counts.sum().item()), algorithmic_bytes=B * P * 5 * 4)
    boxes = [synth_boxes(gen, 1, 119) for _ in range(B)]
    flat = torch.cat(boxes).cuda()
    offs = torch.tensor([0] + list(torch.tensor([b.shape[0] for b in boxes]).cumsum(0)), dtype=torch.int32).cuda()
    out = torch.empty((B, P, 5), device=dev)
    us, n = timed(lambda: ops.ssd_grid_encode(flat, offs, (60, 30, 15, 7), 480, 480, out))
    emit(row=11, what="fd_ssd_grid_encode batch 128 (targets of SSD training)", us=us, launches=n,
         algorithmic_bytes=B * P * 5 * 4, gbps=B * P * 20 / us / 1e3)


if __name__ == "__main__":
    main()
