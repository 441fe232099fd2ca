"""Timestamps of CTA 0 of the stem forward (FD_STEM_TIMING=1)."""
import ctypes
import importlib
import os
import sys

os.environ["FD_STEM_TIMING"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cache = len(sys.argv) > 2 and sys.argv[2] == "cache"
x = torch.rand(B, 3, 480, 480, device="cuda")
w = torch.randn(64, 3, 10, 10, device="cuda") * 0.05
b = torch.randn(64, device="cuda")
y = torch.empty(B, 60, 60, 64, dtype=torch.bfloat16, device="cuda")
xc = torch.zeros(ops.stem_cache_elems(B, 3, 480, 480, 64, 10, 8, 2), dtype=torch.bfloat16, device="cuda") if cache else None
for _ in range(3):
    ops.stem_fwd(x, w, b, y, 8, 2, x_cache=xc)
torch.cuda.synchronize()
n = 32 * 16
buf = (ctypes.c_ulonglong * n)()
L = fd.native.lib()
L.fd_debug_stem_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert L.fd_debug_stem_timing(buf, n) == 0
names = ["cv_wait_stg", "cv_stg_ok", "cv_grp_ok", "cv_chunk0_end", "cv_task_end", "mma_start", "mma_grp0_ok", "mma_issued",
         "epi_acc_ok", "epi_end", "c1_regs", "c1_tma", "c1_stored", "c1_fenced", "c1_grp_ok", "c1_start"]
t00 = buf[1 * 16 + 0]
print("task " + " ".join(f"{s:>13s}" for s in names))
for it in range(1, 9):
    print(f"{it:4d} " + " ".join(f"{buf[it * 16 + k] - t00:13d}" for k in range(16)))
