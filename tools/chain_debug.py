"""Debug driver for the chain kernel: runs one forward chain and reports the CUDA status."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
B, H, W, C, nb = 64, 15, 15, 64, int(os.environ.get("NB", "2"))
g = torch.Generator().manual_seed(3)
x = (torch.randn(B, H, W, C, generator=g) * 0.5).cuda().bfloat16()
w = (torch.randn(2 * nb, C, C, 3, 3, generator=g) * 0.05).cuda()
bias = (torch.randn(2 * nb, C, generator=g) * 0.1).cuda()
wf = torch.empty(2 * nb, 9, C, C, dtype=torch.bfloat16, device="cuda")
ops.pack_conv3x3(w, wf, None)
torch.cuda.synchronize()
outs = [torch.zeros_like(x) for _ in range(nb)]
blocks = []
for k in range(nb):
    d = {"bias1": bias[2 * k], "bias2": bias[2 * k + 1]}
    if mode == "train" or k == nb - 1:
        d["out"] = outs[k]
    if mode == "train":
        d["a"] = torch.zeros_like(x)
        d["mask_a"] = torch.zeros(B, H, W, 2, dtype=torch.int32, device="cuda")
        d["mask_b"] = torch.zeros(B, H, W, 2, dtype=torch.int32, device="cuda")
    blocks.append(d)
ops.resblock_chain_fwd(x, wf, blocks)
try:
    torch.cuda.synchronize()
    print(mode, "dbg", os.environ.get("FD_CHAIN_DBG"), "chain ok; out abs sum", outs[-1].float().abs().sum().item())
except Exception as e:  # noqa: BLE001
    print(mode, "dbg", os.environ.get("FD_CHAIN_DBG"), "FAILED:", str(e)[:200])
    sys.exit(0)
cur = x
for k in range(nb):
    a, s = torch.empty_like(x), torch.empty_like(x)
    ops.conv3x3(cur, wf[2 * k], bias=bias[2 * k], lrelu=True, out=a)
    ops.conv3x3(a, wf[2 * k + 1], bias=bias[2 * k + 1], lrelu=True, residual=cur, out=s)
    cur = s
torch.cuda.synchronize()
print("ref abs sum", cur.float().abs().sum().item(), "equal:", torch.equal(cur, outs[-1]),
      "max diff", (cur.float() - outs[-1].float()).abs().max().item())

if os.environ.get("FD_CHAIN_TIMING"):
    import ctypes
    import numpy as np
    nl = 2 * nb
    buf = (ctypes.c_ulonglong * (2 * 24 * 16))()
    fd.native.lib().fd_debug_chain_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    fd.native.lib().fd_debug_chain_timing(buf, 2 * 24 * 16)
    t = np.array(buf[:], dtype=np.int64).reshape(2, 24, 16)
    t0 = t[0, 0, 0]
    names = {4: "woke", 0: "mma_go", 1: "mma_issued", 5: "mma_done", 2: "epi_top", 6: "waited", 7: "topbar", 8: "prepoll", 3: "acc", 11: "math_done", 13: "bar", 14: "arrived"}
    for l in range(nl):
        for r in range(2):
            print(l, "rank", r, " ".join(f"{names[k]}={t[r, l, k] - t0}" for k in (0, 1, 5, 2, 6, 7, 8, 3, 11, 13)))
