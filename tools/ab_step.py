"""A/B of engine switches on the headline train step (PoolResnet-medium, batch 64, graph replay): python tools/ab_step.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
from bench import synth_batch
dev = torch.device("cuda")
x_cpu, boxes = synth_batch(64)
x = x_cpu.to(dev)
gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, 10, (480, 480), device=dev)


def measure(**switches):
    torch.manual_seed(2)
    m = fd.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10).to(dev).train()
    eng = m.engine
    for k, v in switches.items():
        setattr(eng, k, v)
    eng.bind(dict(m.named_parameters()))
    opt = fd.optim.FlatAdam(eng, lr=1e-4, capturable=True)
    g, pl, n = eng.capture_train_step(x, gt, dropout=True, optimizer=opt)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 200)
    return best, n


for rep in range(2):
    for sw in ({"fuse_pool": True}, {"fuse_pool": False}):
        ms, n = measure(**sw)
        print(sw, f"{ms * 1e3:.1f} us/step, {n} launches", flush=True)
