"""Per-phase clock64 timestamps of CTA 0 of the separable-block kernel (FD_SEP_TIMING=1)."""
import ctypes
import importlib
import os
import sys

os.environ["FD_SEP_TIMING"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
B, H = int(sys.argv[1]), int(sys.argv[2])
pool = len(sys.argv) > 3 and sys.argv[3] == "pool"
x = torch.randn(B, H, H, 64, device="cuda").bfloat16()
w_pw = (torch.randn(2, 64, 64, device="cuda") * 0.1).bfloat16()
w_dw = torch.randn(9, 64, device="cuda") * 0.3
out = torch.empty((B, H // 2, H // 2, 64) if pool else (B, H, H, 64), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.sepblock_fwd(x, w_pw[0], w_dw, w_pw[1], 0.2, out, pool=pool)
torch.cuda.synchronize()
n = 64 * 8
buf = (ctypes.c_ulonglong * n)()
L = fd.native.lib()
L.fd_debug_sep_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert L.fd_debug_sep_timing(buf, n) == 0
names = ["start", "x_tile_ready", "pw1_issued", "acc1_ready", "dw_start", "pw2_phase", "acc2_ready", "epiB_done"]
print("tile", *[f"{s:>14s}" for s in names[1:]], "  next_start")
for it in range(1, 12):
    t = [buf[it * 8 + k] for k in range(8)]
    nxt = buf[(it + 1) * 8]
    print(f"{it:4d}", *[f"{t[k] - t[0]:14d}" for k in range(1, 8)], f"{nxt - t[0]:12d}")
