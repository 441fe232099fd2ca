"""Timing A/B of fd_conv3x3_wgrad_wide against the four fd_conv3x3_wgrad_multi launches it replaces.  python tools/wgrad_wide_debug.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
dev = "cuda"
n3 = 9 * 64 * 64


def timeit(fn, reps=5):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for (nprob, B, H, W) in [(2, 64, 60, 60), (2, 64, 30, 30), (16, 64, 15, 15), (2, 16, 60, 60)]:
    xp = [torch.randn(nprob, B, H, W, 64, device=dev).bfloat16() for _ in range(2)]
    gp = [torch.randn(nprob, B, H, W, 64, device=dev).bfloat16() for _ in range(2)]
    dwp = torch.zeros(nprob * 4 * n3, device=dev)
    db = torch.zeros(nprob * 128, device=dev)
    sub_off = [(c * 2 + r) * n3 for r in range(2) for c in range(2)]
    t_wide = timeit(lambda: ops.conv3x3_wgrad_wide(xp[0], xp[1], gp[0], gp[1], dwp, sub_off, dw_stride=4 * n3, dbias0=db, dbias1=db[64:],
                                                   dbias_stride=128))

    def four():
        for g in range(2):
            for h in range(2):
                ops.conv3x3_wgrad_multi(xp[h], gp[g], dwp[(g * 2 + h) * n3:], 4 * n3, db[g * 64:] if h == 0 else None, 128)
    t_four = timeit(four)
    fl = 2.0 * nprob * B * H * W * 9 * 128 * 128
    print(f"nprob={nprob} B={B} {H}x{W}: wide (2 passes) {t_wide:.1f} us = {fl / t_wide / 1e6:.0f} TF; four 64-channel launches {t_four:.1f} us = {fl / t_four / 1e6:.0f} TF")
