"""Warm per-kernel durations of one SSD(filters=16) train step (SSDEngine, 16 images), eager; FD_NO_PDL=1 recommended."""
import collections, importlib, os, re, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
dev = torch.device("cuda", 0)
torch.manual_seed(2)
m = fd.models.SSD.SSD(filters=16, input_shape=(3, 480, 480)).to(dev).train()
m.engine.bind(dict(m.named_parameters()))
x_cpu, boxes = bench.synth_batch(16, seed_img=10, seed_box=11, kmin=1, kmax=119)
gt = fd.datasets.WIDERFace.dataset_ssd.convert_bbx_to_feature_maps_batch(boxes, (480, 480), device=dev)
x = x_cpu.to(dev)
priors, mult = m._device_priors(dev)
step = lambda: m.engine.train_step(x, gt, priors, mult, 10, dropout=True)
for _ in range(3):
    step()
torch.cuda.synchronize()
R = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(R):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        mm = re.search(r"(\w+)(<[^(]*>)?\(", ev.name.replace("(anonymous namespace)::", ""))
        name = (mm.group(1) + (mm.group(2) or ""))[:48] if mm else ev.name[:48]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(a[1] for a in agg.values())
print(f"sum of kernel durations per step: {tot / R:.1f} us")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{k:50s} n/step={a[0] / R:6.1f} us/step={a[1] / R:8.1f} avg={a[1] / a[0]:7.1f}")
