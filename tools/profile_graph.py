"""Warm, in-graph per-kernel durations (torch.profiler / CUPTI around CUDA-graph replays, programmatic dependent launch off
so that a kernel's duration is its own) of the secondary legs:  python tools/profile_graph.py mbv3|sep|sep128|ssd|f128"""
import os
os.environ.setdefault("FD_NO_PDL", "1")
import collections, importlib, re, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "mbv3"
torch.manual_seed(2)
if which == "mbv3":
    B = 256
    m = fd.models.MobilenetV3Backbone.MobilenetV3Backbone(576, (3, 480, 480), 15).to(dev).eval()
    x = torch.rand(B, 3, 480, 480, device=dev)
    fn = lambda: m(x)
elif which in ("sep", "sep128"):
    B = 256
    m = fd.models.SeparableCNN.SeparableCNN(filters=64 if which == "sep" else 128, input_shape=(3, 480, 480)).to(dev).eval()
    m.engine.bind(dict(m.named_parameters()))
    x = torch.rand(B, 3, 480, 480, device=dev)
    fn = lambda: m.engine.forward(x)
elif which == "ssd":
    B = 16
    m = fd.models.SSD.SSD(filters=16, input_shape=(3, 480, 480)).to(dev).train()
    m.engine.bind(dict(m.named_parameters()))
    x_cpu, boxes = bench.synth_batch(B, kmin=1, kmax=119)
    gt = fd.datasets.WIDERFace.dataset_ssd.convert_bbx_to_feature_maps_batch(boxes, (480, 480), device=dev)
    x = x_cpu.to(dev)
    priors, mult = m._device_priors(dev)
    fn = lambda: m.engine.train_step(x, gt, priors, mult, 10, dropout=True)
else:
    B = 64
    m = fd.models.PoolResnet.PoolResnet(filters=128, input_shape=(3, 480, 480), num_of_patches=10).to(dev).train()
    m.engine.bind(dict(m.named_parameters()))
    x_cpu, boxes = bench.synth_batch(B)
    gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, 10, (480, 480), device=dev)
    x = x_cpu.to(dev)
    fn = lambda: m.engine.train_step(x, gt, dropout=True)
with torch.no_grad():
    g, _, n = bench.capture(fn)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
R = 5
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(R):
    g.replay()
b.record(); torch.cuda.synchronize()
print(f"{which}: batch {B}, {n} launches, {a.elapsed_time(b) / R * 1e3:.1f} us per replay (no PDL)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(R):
        g.replay()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        mm = re.search(r"(\w+)(<[^(]*>)?\(", ev.name.replace("(anonymous namespace)::", ""))
        name = (mm.group(1) + (mm.group(2) or ""))[:48] if mm else ev.name[:48]
        e = agg.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(e[1] for e in agg.values())
print(f"sum of kernel durations per replay: {tot / R:.1f} us")
for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} n={e[0] / R:5.1f} us={e[1] / R:8.1f} avg={e[1] / e[0]:7.1f}")
