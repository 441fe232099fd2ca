import os
os.environ.setdefault("FD_NO_PDL", "1")   # programmatic dependent launch lets a kernel start (and be timed) while its predecessor runs: per-kernel durations are only meaningful without it
"""Warm, in-graph per-kernel durations of one train step (torch.profiler / CUPTI around graph replays)."""
import collections, importlib, os, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
dev = torch.device("cuda", 0)
torch.manual_seed(2)
model = fd.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10).to(dev).train()
eng = model.engine
eng.bind(dict(model.named_parameters()))
x_cpu, boxes = bench.synth_batch(64)
gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, 10, (480, 480), device=dev)
x = x_cpu.to(dev)
graph, pl, n = eng.capture_train_step(x, gt, dropout=True)
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
R = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(R):
        graph.replay()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        import re
        m = re.search(r"(\w+)(<[^(]*>)?\(", ev.name.replace("(anonymous namespace)::", ""))
        name = (m.group(1) + (m.group(2) or ""))[:48] if m else ev.name[:48]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(a[1] for a in agg.values())
print(f"sum of kernel durations per step: {tot / R:.1f} us")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} n/step={a[0] / R:5.1f} us/step={a[1] / R:8.1f} avg={a[1] / a[0]:7.1f}")
