"""Per-kernel device times of the DATA-PARALLEL train step (torch.profiler around graph replays) on rank 0, plus the step
time with and without the exchange.  torchrun --nproc-per-node N tools/profile_step_dp.py"""
import collections, importlib, os, re, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
par = fd.parallel
rank, world, local = par.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import torch.distributed as dist
torch.manual_seed(2)
model = fd.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10).to(dev).train()
eng = model.engine
eng.bind(dict(model.named_parameters()))
par.broadcast_flat(eng.pflat)
x_cpu, boxes = bench.synth_batch(64, seed_img=rank * 2, seed_box=rank * 2 + 1)
gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, 10, (480, 480), device=dev)
x = x_cpu.to(dev)
opt = fd.optim.FlatAdam(eng, lr=1e-4, capturable=True)
split = par.SplitAllReduce.create(eng, eng.plan(64, True), dev) if world > 1 else None


def timeit(g, n=200):
    for _ in range(20):
        g.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


g0, pl, n0 = eng.capture_train_step(x, gt, dropout=True, allreduce=None, optimizer=opt)
t0 = timeit(g0)
g1, pl, n1 = eng.capture_train_step(x, gt, dropout=True, allreduce=split, optimizer=opt)
t1 = timeit(g1)
if rank == 0:
    print(f"world {world}: step without exchange {t0:.1f} us ({n0} launches), with exchange {t1:.1f} us ({n1} launches)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        g1.replay()
    torch.cuda.synchronize()
if rank == 0:
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            mm = re.search(r"(\w+)(<[^(]*>)?\(", ev.name.replace("(anonymous namespace)::", ""))
            name = (mm.group(1) + (mm.group(2) or ""))[:48] if mm else ev.name[:48]
            e = agg.setdefault(name, [0, 0.0])
            e[0] += 1
            e[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"{k:50s} n={e[0] / 10:5.1f} us={e[1] / 10:8.1f} avg={e[1] / e[0]:7.1f}")
if split is not None and split.peer:
    split.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
