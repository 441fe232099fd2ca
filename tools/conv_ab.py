"""A/B of the two fd_conv3x3 kernels (default tap-row fused N=192 vs FD_CONV_ONE_TAP) on the shapes of the train step:
CUDA-graph of `reps` launches over rotating buffers (> L2), CUDA events.  python tools/conv_ab.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
dev = "cuda"
C = 64
w = torch.randn(C, C, 3, 3, device=dev) * 0.05
wf = torch.empty(9, C, C, dtype=torch.bfloat16, device=dev); wd = torch.empty_like(wf)
ops.pack_conv3x3(w, wf, wd)
bias = torch.randn(C, device=dev)
for (B, H, W) in [(64, 60, 60), (64, 30, 30), (64, 15, 15), (16, 240, 240), (16, 120, 120), (16, 60, 60), (16, 30, 30), (16, 15, 15)]:
    nbuf = max(2, int(300e6 // (B * H * W * C * 2 * 3)) + 1)
    xs = [torch.randn(B, H, W, C, device=dev).bfloat16() for _ in range(nbuf)]
    rs = [torch.randn(B, H, W, C, device=dev).bfloat16() for _ in range(nbuf)]
    os_ = [torch.empty_like(xs[0]) for _ in range(nbuf)]
    mo = torch.zeros(B, H, W, 2, dtype=torch.int32, device=dev)
    line = f"B={B:3d} {H:3d}x{W:<3d}"
    for name, fl in (("fused", 0), ("one_tap", ops.CONV_ONE_TAP)):
        def run():
            for i in range(nbuf):
                ops.conv3x3(xs[i], wf, bias=bias, lrelu=True, residual=None if os.environ.get('AB_NO_RES') else rs[i], mask_out=mo, out=os_[i], flags=fl)
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): g.replay()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 10 / nbuf * 1e3
        tf = 2.0 * B * H * W * C * C * 9 / us / 1e6
        line += f" | {name}: {us:7.1f} us {tf:7.1f} TF/s"
        if name == "fused" and os.environ.get("FD_WIDE_TIMING") and os.environ.get("FD_CONV_PAIR"):
            import ctypes
            L = fd.native.lib()
            buf = (ctypes.c_ulonglong * 16)()
            L.fd_debug_wide_timing(buf, 1)
            g.replay(); torch.cuda.synchronize()
            L.fd_debug_wide_timing(buf, 1)
            v = list(buf); tiles = max(1, v[4]); nl = max(1, v[11])
            print(f"    pair kernel, CTA0 per tile (clk): MMA thread total {v[3] / tiles:.0f} = acc_empty {v[0] / tiles:.0f} + input {v[1] / tiles:.0f} + weights {v[2] / tiles:.0f} + issue {(v[3] - v[0] - v[1] - v[2]) / tiles:.0f}; "
                  f"epilogue total {v[7] / tiles:.0f}, wait acc_full {v[5] / tiles:.0f}, wait staging {v[6] / tiles:.0f}; per launch: entry->MMA loop {v[8] / nl:.0f}, MMA end->exit {v[9] / nl:.0f}, tiles {tiles / nl:.1f}")

    print(line, flush=True)
