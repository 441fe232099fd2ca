"""Per-tile timeline of CTA 0 of conv3x3_tc_kernel (FD_CONV_TIMING=1)."""
import ctypes, importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FD_CONV_TIMING"] = "1"
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
B, H, W, C = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[1]) if len(sys.argv) > 1 else 60, 64
mode = sys.argv[2] if len(sys.argv) > 2 else "conv1"
x = torch.randn(B, H, W, C, device="cuda").bfloat16()
res = torch.randn(B, H, W, C, device="cuda").bfloat16()
w = torch.randn(C, C, 3, 3, device="cuda") * 0.05
bias = torch.randn(C, device="cuda")
wf = torch.empty(9, C, C, dtype=torch.bfloat16, device="cuda")
ops.pack_conv3x3(w, wf, None)
out = torch.empty_like(x)
mo = torch.empty(B, H, W, 2, dtype=torch.int32, device="cuda")
for _ in range(3):
    if mode == "conv1":
        ops.conv3x3(x, wf, bias=bias, lrelu=True, mask_out=mo, out=out, flags=ops.CONV_ONE_TAP)
    else:
        ops.conv3x3(x, wf, bias=bias, lrelu=True, residual=res, mask_out=mo, out=out, flags=ops.CONV_ONE_TAP)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    if mode == "conv1":
        ops.conv3x3(x, wf, bias=bias, lrelu=True, mask_out=mo, out=out, flags=ops.CONV_ONE_TAP)
    else:
        ops.conv3x3(x, wf, bias=bias, lrelu=True, residual=res, mask_out=mo, out=out, flags=ops.CONV_ONE_TAP)
e1.record()
torch.cuda.synchronize()
print(mode, H, "avg us (back-to-back, warm L2)", e0.elapsed_time(e1) * 100)
buf = (ctypes.c_ulonglong * 512)()
L = fd.native.lib()
L.fd_debug_conv_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.fd_debug_conv_timing(buf, 512)
t = np.array(buf[:], dtype=np.int64).reshape(64, 8)
t0 = t[0, 0]
names = ["ld_issue", "mma_accfree", "mma_infull", "mma_issued", "epi_start", "epi_math_done", "epi_bar", "epi_store_done"]
for i in range(8):
    print(i, " ".join(f"{names[k]}={t[i, k] - t0}" for k in range(8)))
