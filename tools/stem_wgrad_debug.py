"""A/B timing of stem_wgrad_bf16_kernel with parts switched off (FD_STEM_WG_DBG bits: 1 no MMAs, 2 no bias sums,
4 no drain atomics, 8 no image loads).  Results with a bit set are wrong by construction: timing only."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
dev = torch.device("cuda", 0)
B = 64
x = torch.rand(B, 3, 480, 480, device=dev)
w = torch.randn(64, 3, 10, 10, device=dev) * 0.05
b = torch.zeros(64, device=dev)
y = torch.empty(B, 60, 60, 64, device=dev, dtype=torch.bfloat16)
cache = torch.zeros(ops.stem_cache_elems(B, 3, 480, 480, 64, 10, 8, 2), device=dev, dtype=torch.bfloat16)
ops.stem_fwd(x, w, b, y, 8, 2, x_cache=cache)
g = torch.randn(B, 60, 60, 64, device=dev).to(torch.bfloat16)
dw = torch.zeros_like(w)
db = torch.zeros_like(b)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for mode in (0, 1, 2, 4, 8, 9, 15):
    os.environ["FD_STEM_WG_DBG"] = str(mode)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.stem_wgrad(x, g, dw, db, 8, 2, x_cache=cache)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"dbg={mode:2d}  us: " + " ".join(f"{t:6.1f}" for t in ts[1:]))
