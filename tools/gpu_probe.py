"""GPU probe for the tcgen05 kernels: runs ONE experiment per process (a faulting kernel poisons the
CUDA context) and prints max-abs / relative errors against torch's own conv on the same bf16 inputs.

    python tools/gpu_probe.py conv  <flags> <B> <H> <W>
    python tools/gpu_probe.py wgrad <flags> <B> <H> <W>
"""
import importlib
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main():
    kind, flags, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    torch.manual_seed(0)
    dev = torch.device("cuda")
    C = 64
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    if kind == "conv":
        w = (torch.randn(C, C, 3, 3, device=dev) * 0.05)
        bias = torch.randn(C, device=dev)
        wf = torch.empty(9, C, C, dtype=torch.bfloat16, device=dev)
        wd = torch.empty(9, C, C, dtype=torch.bfloat16, device=dev)
        ops.pack_conv3x3(w, wf, wd)
        out = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=dev)
        ops.conv3x3(x, wf, bias=bias, lrelu=False, out=out, flags=flags)
        torch.cuda.synchronize()
        wq = w.bfloat16().float()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wq, bias, padding=1).permute(0, 2, 3, 1)
        err = (out.float() - ref).abs().max().item()
        print(f"RESULT conv flags={flags} B={B} H={H} W={W} max_abs={err:.4g} rel={rel(out.float(), ref):.4g} "
              f"ref_absmax={ref.abs().max().item():.3g}", flush=True)
        # dgrad packing: conv with wd must equal conv_transpose / input-gradient
        ops.conv3x3(x, wd, out=out, flags=flags)
        torch.cuda.synchronize()
        xin = torch.zeros(B, C, H, W, device=dev, requires_grad=True)
        yy = F.conv2d(xin, wq, None, padding=1)
        (gref,) = torch.autograd.grad(yy, xin, x.float().permute(0, 3, 1, 2))
        gref = gref.permute(0, 2, 3, 1)
        print(f"RESULT dgrad flags={flags} max_abs={(out.float() - gref).abs().max().item():.4g} "
              f"rel={rel(out.float(), gref):.4g}", flush=True)
        # timing
        for _ in range(3):
            ops.conv3x3(x, wf, bias=bias, lrelu=True, out=out, flags=flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            ops.conv3x3(x, wf, bias=bias, lrelu=True, out=out, flags=flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * H * W * C * C * 9
        print(f"TIME conv flags={flags} B={B} H={H} W={W} {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    elif kind == "wgrad":
        g = (torch.randn(B, H, W, C, device=dev) * 0.1).bfloat16()
        dwp = torch.zeros(9, C, C, device=dev)
        db = torch.zeros(C, device=dev)
        ops.conv3x3_wgrad(x, g, dwp, db, flags)
        torch.cuda.synchronize()
        dw = torch.empty(C, C, 3, 3, device=dev)
        ops.unpack_wgrad3x3(dwp, dw)
        wz = torch.zeros(C, C, 3, 3, device=dev, requires_grad=True)
        yy = F.conv2d(x.float().permute(0, 3, 1, 2), wz, None, padding=1)
        (dref,) = torch.autograd.grad(yy, wz, g.float().permute(0, 3, 1, 2))
        bref = g.float().sum(dim=(0, 1, 2))
        print(f"RESULT wgrad flags={flags} B={B} H={H} W={W} max_abs={(dw - dref).abs().max().item():.4g} "
              f"rel={rel(dw, dref):.4g} dbias_rel={rel(db, bref):.4g}", flush=True)
        for _ in range(3):
            ops.conv3x3_wgrad(x, g, dwp, db, flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            ops.conv3x3_wgrad(x, g, dwp, db, flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * H * W * C * C * 9
        print(f"TIME wgrad B={B} H={H} W={W} {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    t = time.time()
    try:
        main()
    except Exception as e:  # noqa: BLE001
        print(f"RESULT {sys.argv[1:]} EXCEPTION {type(e).__name__}: {e}", flush=True)
        sys.exit(1)
    print(f"done in {time.time() - t:.1f}s", flush=True)
