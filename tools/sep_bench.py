"""Timing of the separable-block kernel and of SeparableCNN inference + NMS (BASELINE config 4, batch 256)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops


def timed(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # us


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    torch.manual_seed(0)
    w_pw = (torch.randn(2, 64, 64, device="cuda") * 0.1).bfloat16()
    w_dw = torch.randn(9, 64, device="cuda") * 0.3
    for H in (60, 30, 15):
        # several distinct input buffers so that consecutive launches do not hit the L2 (working set > 126 MB at 60x60)
        n = max(1, min(8, int(400e6 / (B * H * H * 128 * 2))))
        xs = [torch.randn(B, H, H, 64, device="cuda").bfloat16() for _ in range(n)]
        outs = [torch.empty_like(x) for x in xs]

        def run():
            for x, o in zip(xs, outs):
                ops.sepblock_fwd(x, w_pw[0], w_dw, w_pw[1], 0.2, o)
        us = timed(run) / n
        byt = B * H * H * 256
        print(f"sepblock B={B} {H}x{H}: {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s algorithmic ({byt / 1e6:.1f} MB), "
              f"{2 * B * H * H * (2 * 64 * 64 + 9 * 64) / us / 1e6:6.1f} TFLOP/s")
    torch.manual_seed(6)
    m = fd.models.SeparableCNN.SeparableCNN(filters=64, input_shape=(3, 480, 480)).cuda().eval()
    m.engine.bind(dict(m.named_parameters()))
    x = torch.rand(B, 3, 480, 480, device="cuda")
    red = m.reduce_bounding_boxes

    def infer():
        y = m.engine.forward(x)
        red.batch_forward(y)
    n0 = fd.native.launch_count()
    infer()
    print("launches per inference:", fd.native.launch_count() - n0)
    us = timed(infer, reps=10)
    print(f"SeparableCNN(64) infer+NMS B={B}: {us:.1f} us -> {B / us * 1e6:.0f} images/s")
    x8 = (x * 255).to(torch.uint8)

    def infer8():
        y = m.engine.forward(x8)
        red.batch_forward(y)
    us = timed(infer8, reps=10)
    print(f"SeparableCNN(64) infer+NMS B={B} uint8 input: {us:.1f} us -> {B / us * 1e6:.0f} images/s")


if __name__ == "__main__":
    main()
