"""Harness for the round-2 `ncu --set full` capture: one eager PoolResnet-medium (filters=64) train step, one
PoolResnet(filters=128) train step and the config-4 inference forwards between cudaProfilerStart/Stop
(run under `ncu --profile-from-start off`).  Not a benchmark: numbers printed under a profiler are never bench values."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def setup(filters):
    torch.manual_seed(2)
    m = fd.models.PoolResnet.PoolResnet(filters=filters, input_shape=(3, 480, 480), num_of_patches=10).to(dev).train()
    eng = m.engine
    eng.bind(dict(m.named_parameters()))
    x_cpu, boxes = bench.synth_batch(64)
    gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, 10, (480, 480), device=dev)
    return m, eng, x_cpu.to(dev), gt


jobs = []
if which in ("all", "f64"):
    m64, e64, x64, g64 = setup(64)
    jobs.append(lambda: e64.train_step(x64, g64, dropout=True))
if which in ("all", "f128"):
    m128, e128, x128, g128 = setup(128)
    jobs.append(lambda: e128.train_step(x128, g128, dropout=True))
if which in ("all", "infer"):
    xb = torch.rand(256, 3, 480, 480, device=dev)
    sep = fd.models.SeparableCNN.SeparableCNN(filters=64, input_shape=(3, 480, 480)).to(dev).eval()
    sep.engine.bind(dict(sep.named_parameters()))
    jobs.append(lambda: sep.engine.forward(xb))
    mb = fd.models.MobilenetV3Backbone.MobilenetV3Backbone(576, (3, 480, 480), 15).to(dev).eval()
    with torch.no_grad():
        jobs.append(lambda: mb(xb))
for j in jobs:      # warm-up (not profiled)
    with torch.no_grad():
        j()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for j in jobs:
    with torch.no_grad():
        j()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
