import importlib, os, sys, torch
sys.path.insert(0, "/root/repo")
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200"); ops = fd.ops
dev = "cuda"; B = 64
x = torch.rand(B, 3, 480, 480, device=dev); w = torch.randn(128, 3, 10, 10, device=dev) * 0.05; b = torch.zeros(128, device=dev)
ys = [torch.empty(B, 60, 60, 64, device=dev, dtype=torch.bfloat16) for _ in range(2)]
cache = ops.stem_cache(B, (3, 480, 480), (10, 8, 2), dev)
gs = [torch.randn(B, 60, 60, 64, device=dev).bfloat16() for _ in range(2)]
dw = torch.zeros_like(w); db = torch.zeros_like(b)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def t(fn, name):
    ts = []
    for it in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:40s} " + " ".join(f"{v:6.1f}" for v in ts[1:]))
t(lambda: ops.stem_fwd(x, w[:64], b[:64], ys[0], 8, 2, x_cache=cache), "stem_fwd fp32 + cache")
t(lambda: ops.stem_fwd(x, w[64:], b[64:], ys[1], 8, 2), "stem_fwd fp32 (plane 1, old)")
t(lambda: ops.stem_fwd_cached(cache, x.shape, w[64:], b[64:], ys[1], 8, 2), "stem_fwd_cached (plane 1, new)")
t(lambda: ops.stem_wgrad(x, gs[0], dw[:64], db[:64], 8, 2, x_cache=cache), "stem_wgrad one plane")
t(lambda: ops.stem_wgrad_pair(cache, x.shape, gs[0], gs[1], dw, db, 8, 2), "stem_wgrad_pair")
for mode in (1, 4, 8, 9, 13):        # timing switches of the cached kernels: 1 no MMAs, 4 no stores / atomics, 8 no image loads
    os.environ["FD_STEM_WG_DBG"] = str(mode)
    t(lambda: ops.stem_fwd_cached(cache, x.shape, w[64:], b[64:], ys[1], 8, 2), f"stem_fwd_cached dbg={mode}")
os.environ["FD_STEM_WG_DBG"] = "0"
