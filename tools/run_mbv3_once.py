"""One warm-up + N forward passes of the MobilenetV3 backbone at batch B (ncu target).  python tools/run_mbv3_once.py [B] [N]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(9)
m = fd.models.MobilenetV3Backbone.MobilenetV3Backbone(576, (3, 480, 480), 15).cuda().eval()
x = torch.rand(B, 3, 480, 480, device="cuda")
with torch.no_grad():
    m(x[:2])
    for _ in range(N + 1):
        y = m.engine.forward(x)
torch.cuda.synchronize()
print("ok", float(y.mean()))
