"""Timeline of CTA 0 of wgrad3x3_tc_kernel (FD_WGRAD_TIMING=1) + graph-timed duration."""
import ctypes, importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FD_WGRAD_TIMING"] = "1"
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops
H = int(sys.argv[1]) if len(sys.argv) > 1 else 60
nprob = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, C = 64, 64
x = torch.randn(nprob, B, H, H, C, device="cuda").bfloat16()
g = torch.randn(nprob, B, H, H, C, device="cuda").bfloat16()
dw = torch.zeros(nprob, 9 * C * C, device="cuda")
db = torch.zeros(nprob, C, device="cuda")
def run():
    if nprob == 1:
        ops.conv3x3_wgrad(x[0], g[0], dw[0], db[0])
    else:
        ops.conv3x3_wgrad_multi(x, g, dw.view(-1), 9 * C * C, db.view(-1), C)
for _ in range(3):
    run()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(10):
        run()
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print(f"wgrad H={H} nprob={nprob}: {e0.elapsed_time(e1) * 100:.1f} us per launch (graph, warm L2)")
buf = (ctypes.c_ulonglong * 16)()
L = fd.native.lib()
L.fd_debug_wgrad_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.fd_debug_wgrad_timing(buf, 16)
t = np.array(buf[:6], dtype=np.int64)
print("start=0 setup_done=%d epi_wait_acc=%d acc_ready=%d drained=%d end=%d" % tuple(t[1:6] - t[0]))
