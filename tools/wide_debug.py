"""Debug / A-B tool for fd_conv3x3_wide: per-image errors of the epilogue variants, and timings.  python tools/wide_debug.py [B H W gin]"""
import importlib, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
dev = "cuda"
B, H, W, gin = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (150, 15, 15, 2)))
torch.manual_seed(1)
Cin = 64 * gin
x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
res = torch.randn(B, H, W, 128, device=dev).bfloat16()
w = torch.randn(128, Cin, 3, 3, device=dev) * 0.05
bias = torch.randn(128, device=dev)
cs = (torch.rand(B, 128, device=dev) < 0.75).float() / 0.75
wf = torch.empty(1, 1, gin, 9, 128, 64, dtype=torch.bfloat16, device=dev)
ops.pack_conv3x3_wide(w, wf, None)
pl = lambda t: [t[..., 64 * g:64 * (g + 1)].contiguous() for g in range(t.shape[-1] // 64)]
xp = pl(x)
conv = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), bias, padding=1).permute(0, 2, 3, 1)
act = F.leaky_relu(conv, 0.2)
csp = [cs[:, :64].contiguous(), cs[:, 64:].contiguous()]


def report(name, got, ref):
    if os.environ.get("SYNC"):
        torch.cuda.synchronize()
    err = (torch.cat([g.float() for g in got], -1) - ref).abs().flatten(1).max(1).values
    bad = (~(err < 0.05 * max(1.0, ref.abs().max().item()))).nonzero().flatten().tolist()
    print(f"{name:28s} max err {err.max().item():.4g}  bad images: {bad[:20]}{' ...' if len(bad) > 20 else ''} ({len(bad)})")


for rep in range(3):
    out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ops.conv3x3_wide(xp, wf[0, 0], bias=bias, lrelu=True, out=out)
    report("bias+lrelu", out, act)
    out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ops.conv3x3_wide(xp, wf[0, 0], bias=bias, lrelu=True, chan_scale=csp, out=out)
    report("+chan_scale", out, act * cs[:, None, None, :])
    out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ops.conv3x3_wide(xp, wf[0, 0], bias=bias, lrelu=True, residual=pl(res), out=out)
    report("+residual", out, act + res.float())
    out = [torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    mo = [torch.zeros(B, H, W, 2, dtype=torch.int32, device=dev) for _ in range(2)]
    ops.conv3x3_wide(xp, wf[0, 0], bias=bias, lrelu=True, mask_out=mo, out=out)
    report("+mask_out", out, act)
torch.cuda.synchronize()

# timing: rotating buffers > L2, CUDA graph of `nbuf` launches
nbuf = max(2, int(400e6 // (B * H * W * (Cin + 256) * 2)) + 1)
xs = [[torch.randn(B, H, W, 64, device=dev).bfloat16() for _ in range(gin)] for _ in range(nbuf)]
rs = [[torch.randn(B, H, W, 64, device=dev).bfloat16() for _ in range(2)] for _ in range(nbuf)]
os_ = [[torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16) for _ in range(2)] for _ in range(nbuf)]
mo = [torch.zeros(B, H, W, 2, dtype=torch.int32, device=dev) for _ in range(2)]


def run():
    for i in range(nbuf):
        ops.conv3x3_wide(xs[i], wf[0, 0], bias=bias, lrelu=True, residual=rs[i], mask_out=mo, out=os_[i])


side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    run()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    g.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * nbuf)
fl = 2.0 * B * H * W * 9 * Cin * 128
if os.environ.get("FD_WIDE_TIMING"):
    import ctypes
    L = fd.native.lib()
    buf = (ctypes.c_ulonglong * 16)()
    L.fd_debug_wide_timing(buf, 1)
    g.replay(); torch.cuda.synchronize()
    L.fd_debug_wide_timing(buf, 1)
    v = list(buf)
    tiles = max(1, v[4])
    print(f"  CTA0 MMA thread per tile (clk): total {v[3] / tiles:.0f} = wait acc_empty {v[0] / tiles:.0f} + wait input {v[1] / tiles:.0f} + wait weights {v[2] / tiles:.0f} + issue/other {(v[3] - v[0] - v[1] - v[2]) / tiles:.0f};"
          f" per launch: entry->after pdl_wait {v[10] / max(1, v[11]):.0f}, entry->MMA loop {v[8] / max(1, v[11]):.0f}, MMA loop end->exit {v[9] / max(1, v[11]):.0f}, tiles/launch {tiles / max(1, v[11]):.1f};"
          f" share-mode timeline from entry (clk/launch): MMA loop end {v[6] / max(1, v[11]):.0f}, epilogue sees accumulators {v[14] / max(1, v[11]):.0f}, epilogue done {v[15] / max(1, v[11]):.0f};"
          f" slowest CTA (max over launches): pdl_wait->exit {v[12]}, entry->exit {v[13]};"
          f" epilogue thread: total {v[7] / tiles:.0f}, wait acc_full {v[5] / tiles:.0f}, wait staging/residual {v[6] / tiles:.0f}  ({tiles} tiles)")
print(f"B={B} {H}x{W} gin={gin}: {us:.1f} us/launch, {fl / us / 1e6:.0f} TFLOP/s (cta_group={os.environ.get('FD_WIDE_CTA_GROUP', '2')})")
