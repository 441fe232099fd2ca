"""Benchmark of the detection hot path (BASELINE.json metric: images/sec of the train step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic WIDERFace-shaped input:
PoolResnet-medium (F=64, S=10, 10 blocks) forward + summed YoloLoss + backward (models/ModelMeta.py:141,
173-176 of the reference), batch 64 per GPU, <=100 boxes/image, Dropout2d active (model.train()), and for
N>1 the all-reduce of the flat fp32 gradient buffer (one NVLink peer-memory kernel, csrc/comm.cu), then the Adam
step (models/ModelMeta.py:104-112) -- all inside one CUDA graph.  Weak scaling: per-GPU work is fixed.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events with the inputs resident in HBM: the block of K steps
is repeated back to back until the timed region is >= 0.5 s (`timed_region`), so that clocks and power are in the
sustained regime and the NVML sampler sees it.  `e2e` runs the same step through the public API
(`model.graphed_train_step`) from pinned HOST buffers (H2D of images + targets and D2H of the loss inside the timed
region).  `--impl reference` times the reference's own CPU implementation of the same step (the real reference modules
from the git-ignored baseline/_ref when present, else the oracle port), all host threads.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "pytorch-face-detection-from-scratch_b200"

B_PER_GPU = 64
S = 10
FLOPS_FWD_PER_IMG = 1069.4e6          # SURVEY 8d, PoolResnet F=64
LR = 1e-4                             # models/ModelMeta.py:86
MIN_TIMED_S = 0.5


def synth_boxes(gen, kmin, kmax, size=480):
    """SURVEY 8d boxes: integer-valued (1,x,y,w,h) f32, log-uniform sizes clipped to the image."""
    k = int(torch.randint(kmin, kmax + 1, (1,), generator=gen))
    x = torch.randint(0, size, (k,), generator=gen).float()
    y = torch.randint(0, size, (k,), generator=gen).float()
    lw = torch.rand(k, generator=gen) * (math.log(240.0) - math.log(4.0)) + math.log(4.0)
    lh = torch.rand(k, generator=gen) * (math.log(240.0) - math.log(4.0)) + math.log(4.0)
    w = torch.minimum(torch.round(torch.exp(lw)), size - x).clamp(min=1)
    h = torch.minimum(torch.round(torch.exp(lh)), size - y).clamp(min=1)
    return torch.stack([torch.ones(k), x, y, w, h], dim=1)


def synth_batch(B, seed_img=0, seed_box=1, kmin=1, kmax=100):
    """x = rand(B,3,480,480); K~U{kmin..kmax} integer boxes per image, log-uniform sizes."""
    gx = torch.Generator().manual_seed(seed_img)
    x = torch.rand(B, 3, 480, 480, generator=gx)
    gb = torch.Generator().manual_seed(seed_box)
    boxes = [synth_boxes(gb, kmin, kmax) for _ in range(B)]
    return x, boxes


def seeded_params(filters=64, seed=2):
    """Default-initialised PoolResnet weights in the reference's construction order (models/PoolResnet.py:70-89)."""
    torch.manual_seed(seed)
    p = {}
    c = torch.nn.Conv2d(3, filters, 10, stride=8, padding=2)
    p["conv1.weight"], p["conv1.bias"] = c.weight.detach(), c.bias.detach()
    for b in range(10):
        for n in ("conv1", "conv2"):
            c = torch.nn.Conv2d(filters, filters, 3, padding=1)
            p[f"residual_blocks.{b}.{n}.weight"] = c.weight.detach()
            p[f"residual_blocks.{b}.{n}.bias"] = c.bias.detach()
    c = torch.nn.Conv2d(filters, 5, 6)
    p["out.weight"], p["out.bias"] = c.weight.detach(), c.bias.detach()
    return p


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def _install_reference_stubs():
    import types

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    class _LM(torch.nn.Module):
        def log(self, *a, **k):
            pass

    mod("albumentations")
    mod("albumentations.pytorch")
    mod("albumentations.pytorch.transforms", ToTensorV2=object)
    mod("torchinfo", summary=lambda *a, **k: None)
    mod("ptflops", get_model_complexity_info=lambda *a, **k: (0, 0))
    mod("pytorch_lightning", LightningModule=_LM, Trainer=object, LightningDataModule=object)
    mod("gdown")


def reference_step_factory(Bs):
    """The reference's own train step on CPU (train_model.py:27-39 + models/ModelMeta.py:141,173-176,104-112):
    ``model.train(); y_hat = model(x); loss = sum_i yolo_loss(y_hat[i], y[i]); loss.backward(); Adam.step()``.
    Returns (step, kind, describe): the REAL reference modules from baseline/_ref when that copy is present (5 stub
    modules for packages off the hot path), else the oracle port of the same step."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    x, boxes = synth_batch(Bs)
    if os.path.isfile(os.path.join(ref_root, "models", "PoolResnet.py")):
        _install_reference_stubs()
        sys.path.insert(0, ref_root)
        from datasets.WIDERFace.dataset import WIDERFaceDataset         # the reference's encoder
        from losses.YoloLoss import yolo_loss
        from models.PoolResnet import PoolResnet
        torch.manual_seed(2)
        model = PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=S, num_of_residual_blocks=10).train()
        ds = WIDERFaceDataset(None, S, (3, 480, 480))
        gt = torch.stack([ds.convert_bbx_to_feature_map(b, (480, 480)) for b in boxes])
        opt = torch.optim.Adam(model.parameters(), lr=LR)        # SAMSGD == its base Adam numerically

        def step():
            opt.zero_grad()
            y_hat = model(x)
            loss = 0
            for i in range(Bs):
                loss = loss + yolo_loss(y_hat[i], gt[i])
            loss.backward()
            opt.step()
            return float(loss.detach())

        return step, "reference", "the reference's own modules (baseline/_ref: models/PoolResnet.py, losses/YoloLoss.py), model.train()"
    from oracle import backbone_oracle as bo
    from oracle import yolo_oracle as yo
    gt = torch.stack([torch.from_numpy(yo.grid_encode(b.numpy(), S, 480, 480)) for b in boxes])
    p = seeded_params()
    ost = {}
    gd = torch.Generator().manual_seed(7)

    def step():
        # train-mode Dropout2d like the reference's model.train(): p = 0.25 per block, 0.5 before the head
        scales = [(torch.rand(Bs, 64, 1, 1, generator=gd) < 0.75).float() / 0.75 for _ in range(10)]
        scales.append((torch.rand(Bs, 64, 1, 1, generator=gd) < 0.5).float() / 0.5)
        _, loss, grads = bo.train_step(x, gt, p, S, drop_scales=scales)
        bo.adam_update(p, grads, ost, lr=LR)
        return float(loss)

    return step, "port", "oracle port of the reference step (baseline/_ref absent), train-mode Dropout2d"


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone runs; each step is a bounded 16-image sample of the 64-image batch."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 16
    step, kind, what = reference_step_factory(Bs)
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t0 > 150.0:          # keep the whole run within a few minutes on any host
            break
    dt = time.perf_counter() - t0
    v = Bs * done / dt
    sample = (f"{done} steps x {Bs} images (of the {B_PER_GPU}-image batch), {what}, torch {torch.__version__} CPU fp32, "
              f"{torch.get_num_threads()} threads")
    line = {"impl": "reference", "metric": "train_images_per_sec", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": done, "warmup": args.warmup, "ms_per_step": dt / done * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "PoolResnet-medium (filters=64, S=10, 10 blocks, 480x480) train step: forward + summed "
                        "YoloLoss + backward" + (" + gradient all-reduce" if n > 1 else "") + " + Adam step",
            "global_batch": B_PER_GPU * n, "batch_per_gpu": B_PER_GPU, "boxes_per_image": "1..100",
            "parallelism": f"dp{n}", "dropout": "train-mode Dropout2d",
            "l2": "per-step working set (177 MB fp32 images + ~1 GB bf16 activations) exceeds the 126 MB L2; no flush needed"}


def cpu_baseline():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 16
    step, kind, what = reference_step_factory(Bs)
    step()
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 40):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": Bs * n / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{n} train steps x {Bs} images of the same synthetic workload ({what}; torch CPU fp32)"}


# ------------------------------------------------------------------------------------------------ timing helpers
def timed_graph_region(step, k_steps, warmup, barrier, par, dev):
    """W warm-up steps, then blocks of K steps back to back until the region is >= MIN_TIMED_S; CUDA events on the
    launching stream, barrier + synchronize on both sides, max over ranks.  Returns (ms_per_step, reps, total_ms)."""
    for _ in range(warmup):
        step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k_steps):
        step()
    b.record()
    barrier()
    pilot_ms = par.max_over_ranks(a.elapsed_time(b), dev)
    reps = max(1, min(2000, int(math.ceil(MIN_TIMED_S * 1e3 / max(pilot_ms, 1e-3)))))
    barrier()
    a.record()
    for _ in range(reps * k_steps):
        step()
    b.record()
    barrier()
    total_ms = par.max_over_ranks(a.elapsed_time(b), dev)
    return total_ms / (reps * k_steps), reps, total_ms


def capture(fn, warm=2):
    """Warm up on a side stream, then capture `fn` into a CUDA graph.  Returns (graph, result, launches)."""
    fd = importlib.import_module(PKG)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    n0 = fd.native.launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = fn()
    return g, res, fd.native.launch_count() - n0


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="headline + e2e only (profiling runs)")
    ap.add_argument("--nccl", action="store_true", help="N>1: use the NCCL all-reduce instead of the peer-memory kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    fd = importlib.import_module(PKG)
    import __graft_entry__ as ge
    if rank == 0 or not os.path.exists(fd.native.LIB_PATH):
        ge.build()
    par = fd.parallel
    rank, world, local = par.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    peaks = load_peaks()

    # ---------------- model + synthetic data (each rank its own shard: different seeds)
    torch.manual_seed(2)
    model = fd.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=S).to(dev).train()
    eng = model.engine
    eng.bind(dict(model.named_parameters()))
    par.broadcast_flat(eng.pflat)
    x_cpu, boxes = synth_batch(B_PER_GPU, seed_img=rank * 2, seed_box=rank * 2 + 1)
    gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, S, (480, 480), device=dev)
    x = x_cpu.to(dev)
    B = B_PER_GPU

    # Adam (the reference's optimizer, ModelMeta.py:104-112) is part of the step: one kernel over the flat buffers,
    # step count on the device so that it replays from the graph.  N > 1: the gradient all-reduce is ONE peer-memory
    # kernel inside the same graph (csrc/comm.cu); NCCL only if peer memory cannot be mapped on this box.
    opt = fd.optim.FlatAdam(eng, lr=LR, capturable=True)
    # N > 1: the gradient exchange is TWO peer-memory all-reduce kernels inside the step's graph (csrc/comm.cu,
    # parallel.SplitAllReduce): the packed accumulators of the fused chain on a side stream, overlapped with the rest of
    # the backward pass, and the remainder just before Adam; NCCL only if peer memory cannot be mapped on this box.
    split = None
    if world > 1:
        split = par.SplitAllReduce.create(eng, eng.plan(B, True), dev, use_nccl=args.nccl)
    peer_ar = split if (split is not None and split.peer) else None
    collective = "none" if world == 1 else ("2 nvlink peer-memory all-reduce kernels in the graph (early one overlapped "
                                            "with the backward pass)" if peer_ar else "nccl all_reduce")
    eager_ar = peer_ar if peer_ar is not None else (par.allreduce_grads if world > 1 else None)

    def eager_step():
        return eng.train_step(x, gt, dropout=True, allreduce=eager_ar, optimizer=opt)

    graph = None
    if args.no_graph:
        n0 = fd.native.launch_count()
        pl = eager_step()
        per_step_launches = fd.native.launch_count() - n0
        step = eager_step
    elif world > 1 and peer_ar is None:
        graph, pl, per_step_launches = eng.capture_train_step(x, gt, dropout=True)
        per_step_launches += 1

        def step():
            graph.replay()
            par.allreduce_grads(eng.gflat)
            opt.step()
            return pl
    else:
        graph, pl, per_step_launches = eng.capture_train_step(x, gt, dropout=True, allreduce=peer_ar, optimizer=opt)

        def step():
            graph.replay()
            return pl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    ms_step, reps, total_ms = timed_graph_region(step, args.steps, args.warmup, barrier, par, dev)
    sampler.stop_flag = True
    loss_val = float(pl.loss.sum().item())
    value = world * B / (ms_step * 1e-3)
    comm_status = peer_ar.status() if peer_ar is not None else 0        # 0 = every exchange completed

    # ---------------- e2e: the public graphed-step API, pinned host inputs, H2D + D2H inside the timed region
    def e2e_leg(host_images):
        gth = gt.cpu().pin_memory()
        xh = host_images.pin_memory()
        gstep = model.graphed_train_step(B, host_images.dtype, optimizer=opt, allreduce=peer_ar if peer_ar else None,
                                         n_buffers=2)
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i])            # the step that last read buffer i has finished
                gstep.x[i].copy_(xh, non_blocking=True)
                gstep.gt[i].copy_(gth, non_blocking=True)
                ready[i].record(copy_stream)

        def run(nsteps):
            prefetch(0)
            for it in range(nsteps):
                i = it & 1
                torch.cuda.current_stream().wait_event(ready[i])
                if it + 1 < nsteps:
                    prefetch(i ^ 1)                  # the next batch's H2D overlaps this step's compute
                loss = gstep.replay(i)               # forward + loss + backward (+ all-reduce) + Adam: one graph launch
                consumed[i].record()
                if world > 1 and peer_ar is None:
                    par.allreduce_grads(eng.gflat)
                _ = loss.sum().item()                # D2H of the step's result

        for ev in consumed:
            ev.record()
        run(3)
        barrier()
        n_e2e = max(5, min(args.steps, 20))
        # as many K-step blocks as fit ~0.5 s
        t0 = time.perf_counter()
        run(n_e2e)
        barrier()
        pilot = par.max_over_ranks(time.perf_counter() - t0, dev)
        reps_e = max(1, min(50, int(math.ceil(MIN_TIMED_S / max(pilot, 1e-4)))))
        barrier()
        t0 = time.perf_counter()
        run(n_e2e * reps_e)
        barrier()
        secs = par.max_over_ranks(time.perf_counter() - t0, dev)
        return {"value": world * B * n_e2e * reps_e / secs, "unit": "images/s",
                "h2d_bytes_per_step": xh.numel() * xh.element_size() + gth.numel() * 4, "d2h_bytes_per_step": 4 * B,
                "steps": n_e2e * reps_e, "ms_per_step": secs / (n_e2e * reps_e) * 1e3}

    e2e = e2e_leg(x_cpu)
    e2e["note"] = ("fp32 images from pinned host memory (the reference's training input type), double-buffered H2D "
                   "overlapped with compute; public API model.graphed_train_step (one CUDA-graph launch per step); "
                   "bounded by the 177 MB/step host->device copy (PCIe)")
    e2e_u8 = e2e_leg((x_cpu * 255.0).round().to(torch.uint8))
    e2e_u8["note"] = ("same step fed with the uint8 images the reference's dataset holds before `img / 255` "
                      "(datasets/WIDERFace/dataset.py:146); the division is fused into the stem kernel")

    extra = {}
    if not args.no_extra_legs:
        # ---------------- inference: eval forward + batched decode + NMS (BASELINE metric's second half)
        model.eval()
        red = model.reduce_bounding_boxes

        def infer_fn():
            pli = eng.forward(x, train=False, dropout=False)
            return red.batch_forward(pli.y)

        g_inf, (inf_boxes, inf_counts), inf_launches = capture(infer_fn)
        inf_ms, _, _ = timed_graph_region(g_inf.replay, args.steps, args.warmup, barrier, par, dev)
        extra["infer"] = {"metric": "infer_nms_images_per_sec", "value": world * B / (inf_ms * 1e-3),
                          "unit": "images/s", "ms_per_step": inf_ms, "launches_per_step": inf_launches,
                          "workload": "PoolResnet-medium eval forward + YOLO decode + score threshold (0.5) + IoU NMS "
                                      "(0.5), batch 64 per GPU, inputs resident in HBM, CUDA graph",
                          "kept_boxes_batch": int(inf_counts.sum().item())}
        model.train()
        for name, fn in (("infer_separable", separable_infer), ("infer_mobilenet", mobilenet_infer),
                         ("train_ssd", ssd_train), ("train_f128", f128_train)):
            try:      # a secondary leg must not lose the headline line
                extra[name] = fn(fd, dev, world, args, barrier, par, peaks)
            except Exception as exc:  # noqa: BLE001
                extra[name] = {"error": repr(exc)}

    roof = cpu_base = lib_base = None
    if rank == 0:
        roof = conv_roofline(fd, eng, pl, dev, peaks, total_ms * 1e-3)
        if world == 1 and not args.no_extra_legs:
            try:
                lib_base = gpu_library_baseline(dev, x, gt)
            except Exception as exc:  # noqa: BLE001
                lib_base = {"error": repr(exc)}
    if peer_ar is not None:
        peer_ar.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()          # non-zero ranks leave now: nothing spins while rank 0 does CPU work
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_baseline()          # N = 1 only: at N > 1 the host cores are shared with the other ranks
        line = {"metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(world), "loss": loss_val, "collective": collective, "comm_status": comm_status,
                "timed_region": {"blocks_of_k_steps": reps, "steps": reps * args.steps, "seconds": total_ms * 1e-3,
                                 "why": "the K-step block is repeated until the region is >= 0.5 s"},
                "optimizer_steps": opt.device_steps(),
                "clocks": sampler.summary(),
                "e2e": e2e, "e2e_u8": e2e_u8,
                "gpu_launches": per_step_launches * reps * args.steps,
                "launches_per_step": per_step_launches, "cuda_graph": graph is not None,
                "achieved_tflops_step": 3 * FLOPS_FWD_PER_IMG * B / (ms_step * 1e-3) / 1e12,
                "roofline": roof, "cpu_baseline": cpu_base, "gpu_library_baseline": lib_base}
        line.update(extra)
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ secondary legs
def _hbm_roofline(kernel, byt, us, peak, note=None):
    r = {"kernel": kernel, "bound": "hbm", "achieved": byt / us / 1e3, "peak": peak, "unit": "GB/s",
         "frac": byt / us / 1e3 / peak, "avg_launch_us": us, "algorithmic_bytes_per_launch": byt, "traffic": None}
    if note:
        r["note"] = note
    return r


def _time_graph(fn, reps=10):
    g, _, _ = capture(fn, warm=1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3          # us per replay


def separable_infer(fd, dev, world, args, barrier, par, peaks):
    """SeparableCNN(filters=64) eval forward (stem, 10 fused separable blocks with fused pooling, head) + batched
    decode + NMS, batch 256 per GPU, fp32 images resident in HBM, one CUDA graph; plus the HBM roofline of the
    dominant kernel of that path (fd_sepblock_fwd at 60x60: algorithmic bytes = read x + write the pooled y)."""
    Bs = 256
    torch.manual_seed(6)
    m = fd.models.SeparableCNN.SeparableCNN(filters=64, input_shape=(3, 480, 480)).to(dev).eval()
    m.engine.bind(dict(m.named_parameters()))
    xs = torch.rand(Bs, 3, 480, 480, device=dev)
    red = m.reduce_bounding_boxes
    g, _, launches = capture(lambda: red.batch_forward(m.engine.forward(xs)))
    ms, _, _ = timed_graph_region(g.replay, max(5, min(args.steps, 20)), 3, barrier, par, dev)
    out = {"metric": "infer_nms_images_per_sec", "value": world * Bs / (ms * 1e-3), "unit": "images/s",
           "ms_per_step": ms, "launches_per_step": launches,
           "workload": "SeparableCNN(filters=64, 10 blocks, 480x480; BASELINE config 4) eval forward + YOLO decode + "
                       "threshold + NMS, batch 256 per GPU, inputs resident in HBM, CUDA graph"}
    peak = peaks.get("hbm_gbs", 6436.4)
    w_pw = (torch.randn(2, 64, 64, device=dev) * 0.1).bfloat16()
    w_dw = torch.randn(9, 64, device=dev) * 0.3
    bufs = [torch.randn(Bs, 60, 60, 64, device=dev).bfloat16() for _ in range(3)]      # rotating: larger than the L2
    outs = [torch.empty((Bs, 30, 30, 64), dtype=torch.bfloat16, device=dev) for _ in range(3)]

    def run():
        for t, o in zip(bufs, outs):
            fd.ops.sepblock_fwd(t, w_pw[0], w_dw, w_pw[1], 0.2, o, pool=True)

    us = _time_graph(run) / len(bufs)
    out["roofline"] = _hbm_roofline("sepblock_fwd_kernel (60x60 block, MaxPool2d fused)", Bs * 3600 * 128 + Bs * 900 * 128,
                                    us, peak, "instruction-issue bound (two tcgen05 GEMM epilogues + the CUDA-core "
                                    "depthwise stage per tile), not memory bound: DESIGN.md 4.7")
    tr = _traffic("r2_traffic.json", "sepblock_fwd_kernel")
    if tr is not None:
        out["roofline"]["traffic"] = tr
    return out


def _traffic(fname, kernel):
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", fname)))
        return tj[kernel]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None


def mobilenet_infer(fd, dev, world, args, barrier, par, peaks):
    """BASELINE config 4, first model: MobilenetV3Backbone (tf_mobilenetv3_small_100 graph, random-init weights with
    non-trivial BatchNorm statistics) eval forward + decode + NMS, batch 256 per GPU, one CUDA graph; HBM roofline of its
    largest pointwise GEMM (conv_pw 16 -> 72 at 120x120: read 32 B + write 144 B per pixel)."""
    Bs = 256
    torch.manual_seed(9)
    m = fd.models.MobilenetV3Backbone.MobilenetV3Backbone(576, (3, 480, 480), 15).to(dev).eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
    xs = torch.rand(Bs, 3, 480, 480, device=dev)
    red = m.reduce_bounding_boxes
    with torch.no_grad():
        m(xs[:2])                                   # prepares (folds + packs) the weights outside the graph
    g, _, launches = capture(lambda: red.batch_forward(m.engine.forward(xs)))
    ms, _, _ = timed_graph_region(g.replay, max(5, min(args.steps, 20)), 3, barrier, par, dev)
    out = {"metric": "infer_nms_images_per_sec", "value": world * Bs / (ms * 1e-3), "unit": "images/s",
           "ms_per_step": ms, "launches_per_step": launches,
           "workload": "MobilenetV3Backbone (timm tf_mobilenetv3_small_100 minus 5 children + 3x3 head, 480x480; BASELINE "
                       "config 4) eval forward + YOLO decode + threshold + NMS, batch 256 per GPU, inputs resident in HBM, "
                       "CUDA graph", "achieved_gflops": 512.5e6 * Bs / (ms * 1e-3) / 1e9}
    peak = peaks.get("hbm_gbs", 6436.4)
    ops = fd.ops
    M, K, N = Bs * 120 * 120, 16, 72
    w = torch.randn(N, K, device=dev) * 0.2
    packed = torch.empty(ops.pw_packed_elems(N, K), dtype=torch.bfloat16, device=dev)
    ops.pw_pack(w, None, packed)
    bpad = torch.zeros(ops.pw_padded_n(N), device=dev)
    xa = [torch.randn(M, K, device=dev).bfloat16() for _ in range(2)]
    ya = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(2)]

    def run():
        for a, o in zip(xa, ya):
            ops.pw_conv(a, packed, bpad, N, ops.ACT_RELU, o)

    us = _time_graph(run) / len(xa)
    out["roofline"] = _hbm_roofline("pw_gemm_kernel (conv_pw 16->72 at 120x120, batch 256)", M * (K + N) * 2, us, peak)
    tr = _traffic("r2_traffic.json", "pw_gemm_kernel")
    if tr is not None:
        out["roofline"]["traffic"] = tr
    return out


def ssd_train(fd, dev, world, args, barrier, par, peaks):
    """BASELINE config 5: SSD(filters=16) train step (forward + ssd_loss(.., 10) + backward), 16 images per GPU
    (128 over 8 GPUs), synthetic targets encoded at the four scales."""
    if not hasattr(fd.models, "SSD"):
        return {"error": "models.SSD not built"}
    return fd.models.SSD.bench_train_step(fd, dev, world, args, barrier, par, timed_graph_region, capture, synth_batch)


def f128_train(fd, dev, world, args, barrier, par, peaks):
    """PoolResnet(filters=128): the width train_model.py:17 trains (SURVEY 8d config C2, F = 128) -- forward + summed
    YoloLoss + backward + Adam, batch 64 per GPU, train-mode dropout, one CUDA graph.  Its 3x3 convolutions run on the
    cta_group::2 kernels (fd_conv3x3_wide forward / input gradient, fd_conv3x3_wgrad_wide); their rooflines are measured
    by re-issuing the step's launches of each kernel back to back in one graph (as conv_roofline does)."""
    if world > 1:
        return {"skipped": "single-GPU leg (the data-parallel exchange is benchmarked on the headline configuration)"}
    ops = fd.ops
    B = B_PER_GPU
    torch.manual_seed(4)
    m = fd.models.PoolResnet.PoolResnet(filters=128, input_shape=(3, 480, 480), num_of_patches=S).to(dev).train()
    eng = m.engine
    eng.bind(dict(m.named_parameters()))
    x_cpu, boxes = synth_batch(B, seed_img=20, seed_box=21)
    gt = fd.datasets.WIDERFace.dataset.convert_bbx_to_feature_map_batch(boxes, S, (480, 480), device=dev)
    x = x_cpu.to(dev)
    topt = m.flat_optimizer(lr=LR, capturable=True)      # fd_adam_flat over the engine's flat buffers: one launch
    topt._ensure_state()
    for name, prm in m.named_parameters():
        prm.grad = eng.grad_view(name)
    calls = {"conv": [], "wgrad": [], "chain": []}
    o_conv, o_wgrad, o_chain = ops.conv3x3_wide, ops.conv3x3_wgrad_wide, ops.conv3x3_wide_chain

    def rec_chain(xs, w, layers, **kw):
        calls["chain"].append((xs, w, layers, kw))
        o_chain(xs, w, layers, **kw)

    def rec_conv(xp, w, **kw):
        calls["conv"].append((xp, w, kw))
        o_conv(xp, w, **kw)

    def rec_wgrad(*a, **kw):
        calls["wgrad"].append((a, kw))
        o_wgrad(*a, **kw)

    ops.conv3x3_wide, ops.conv3x3_wgrad_wide, ops.conv3x3_wide_chain = rec_conv, rec_wgrad, rec_chain
    try:
        eng.train_step(x, gt, dropout=True)
        torch.cuda.synchronize()
    finally:
        ops.conv3x3_wide, ops.conv3x3_wgrad_wide, ops.conv3x3_wide_chain = o_conv, o_wgrad, o_chain
    g, _, launches = capture(lambda: eng.train_step(x, gt, dropout=True, optimizer=topt))
    ms, _, _ = timed_graph_region(g.replay, max(3, min(args.steps, 20)), 3, barrier, par, dev)
    burst = peaks.get("bf16_tflops") or 1590.0

    try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:  # noqa: BLE001
        tj = {}

    def roof(sel, run_one, flops_of, kernel, tkey):
        if not sel:
            return None
        us = _time_graph(lambda: [run_one(c) for c in sel], 10)
        fl = sum(flops_of(c) for c in sel)
        t = tj.get(tkey, {})
        return {"kernel": kernel, "bound": "tensor", "achieved": fl / us / 1e6, "peak": burst, "unit": "TFLOP/s",
                "frac": fl / us / 1e6 / burst, "launches": len(sel), "avg_launch_us": us / len(sel),
                "traffic": t.get("dram_bytes_per_launch"), "traffic_unit": "bytes/launch", "traffic_source": t.get("source"),
                "tensor_pipe_active_pct_ncu": t.get("tensor_pipe_active_pct")}

    big = [c for c in calls["conv"] if c[0][0].shape[1] == 60]
    conv_fl = lambda c: 2.0 * c[0][0].shape[0] * c[0][0].shape[1] * c[0][0].shape[2] * 9 * 64 * len(c[0]) * 128
    wg_fl = lambda c: 2.0 * c[0][0].numel() / 64 * 9 * 128 * 128
    chain_fl = lambda c: 2.0 * len(c[2]) * c[0][0].shape[1] * c[0][0].shape[2] * c[0][0].shape[3] * 9 * 128 * 128
    n_chain_layers = sum(len(c[2]) for c in calls["chain"])
    return {"metric": "train_images_per_sec", "value": B / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
            "launches_per_step": launches, "cuda_graph": True,
            "achieved_tflops_algorithmic": 3 * 3.997e9 * B / (ms * 1e-3) / 1e12,
            "workload": "PoolResnet(filters=128, S=10, 10 blocks, 480x480; the width train_model.py:17 trains) train step: "
                        "forward + summed YoloLoss + backward + Adam (fd_adam_flat), batch 64, train-mode Dropout2d",
            "roofline_conv_wide_60x60": roof(big, lambda c: o_conv(c[0], c[1], **c[2]), conv_fl,
                                             "conv3x3_wide_kernel<2> (tcgen05.mma.cta_group::2, M=256 N=128; the 60x60 forward + "
                                             "input-gradient launches of one step)", "conv3x3_wide_kernel_60x60"),
            "roofline_conv_wide_all": roof(calls["conv"], lambda c: o_conv(c[0], c[1], **c[2]), conv_fl,
                                           f"conv3x3_wide_kernel<2>, the {len(calls['conv'])} single-layer launches of one step "
                                           "(60x60 and 30x30 maps)", "conv3x3_wide_kernel_all"),
            "roofline_conv_wide_chain": roof(calls["chain"], lambda c: o_chain(c[0], c[1], c[2], **c[3]), chain_fl,
                                             f"conv3x3_wide_chain_kernel: the {n_chain_layers} convolutions on 15x15 maps (8 residual "
                                             f"blocks forward, 8 backward) in {len(calls['chain'])} launches -- one CTA pair per image, "
                                             "64 of 74 pairs busy, each layer's MMAs wait for the previous layer's epilogue",
                                             "conv3x3_wide_chain_kernel"),
            "roofline_wgrad_wide": roof(calls["wgrad"], lambda c: o_wgrad(*c[0], **c[1]), wg_fl,
                                        "wgrad3x3_wide_kernel (cta_group::2, two passes per call)", "wgrad3x3_wide_kernel")}


def gpu_library_baseline(dev, x, gt):
    """The library comparator on the SAME box ("beat cuDNN", SURVEY 2.1 / BASELINE.md 3 step 5): the same train step
    (PoolResnet-medium, batch 64, train-mode Dropout2d, summed YoloLoss vectorised over the batch, Adam) written with
    stock torch modules -- eager cuDNN convolutions, (a) fp32 with TF32 off, (b) bf16 autocast + channels_last.  Not the
    oracle and not the product: plain PyTorch as a user of the reference would run it on this GPU."""
    import torch.nn as nn
    import torch.nn.functional as F

    class Block(nn.Module):
        def __init__(self, f, pool):
            super().__init__()
            self.c1, self.c2, self.pool = nn.Conv2d(f, f, 3, padding=1), nn.Conv2d(f, f, 3, padding=1), pool

        def forward(self, t):
            y = F.leaky_relu(self.c2(F.leaky_relu(self.c1(t), 0.2)), 0.2)
            y = F.dropout2d(y, 0.25, self.training) + t
            return F.max_pool2d(y, 2) if self.pool else y

    class Net(nn.Module):
        def __init__(self, f=64):
            super().__init__()
            self.stem = nn.Conv2d(3, f, 10, stride=8, padding=2)
            self.blocks = nn.Sequential(*[Block(f, k < 2) for k in range(10)])
            self.out = nn.Conv2d(f, 5, 6)

        def forward(self, t):
            return torch.sigmoid(self.out(F.dropout2d(self.blocks(self.stem(t)), 0.5, self.training)))

    def yolo_sum(p, g):            # losses/YoloLoss.py:4-44 vectorised over the batch (sum, ModelMeta.py:173-176)
        g0 = g[:, 0]
        xy = 3 * g0 * ((g[:, 1] - p[:, 2]) ** 2 + (g[:, 2] - p[:, 1]) ** 2)
        wh = 3 * g0 * ((g[:, 3].sqrt() - p[:, 3].sqrt()) ** 2 + (g[:, 4].sqrt() - p[:, 4].sqrt()) ** 2)
        cf = (g0 + (1 - g0) / p.shape[2]) * (g0 - p[:, 0]) ** 2
        return (xy + wh + cf).sum()

    out = {"what": "stock PyTorch (eager, cuDNN) train step of the same model / batch / loss / Adam on this GPU",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    Bq = x.shape[0]
    for name, amp, cl in (("fp32", False, False), ("bf16_autocast_channels_last", True, True)):
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        torch.manual_seed(2)
        net = Net().to(dev).train()
        xi = x
        if cl:
            net = net.to(memory_format=torch.channels_last)
            xi = x.contiguous(memory_format=torch.channels_last)
        opt = torch.optim.Adam(net.parameters(), lr=LR, fused=True)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                y = net(xi)
            loss = yolo_sum(y.float(), gt)
            loss.backward()
            opt.step()

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        a.record()
        for _ in range(n):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        out[name] = {"ms_per_step": ms, "images_per_sec": Bq / (ms * 1e-3)}
        del net, opt
    torch.backends.cudnn.benchmark = False
    return out


def conv_roofline(fd, eng, pl, dev, peaks, timed_seconds):
    """Roofline of the dominant kernel: every conv3x3_tc launch of one forward+backward (same arguments, same
    buffers) is re-issued back to back inside ONE CUDA graph, and the graph is timed with CUDA events on the
    launching stream -- device time of the kernel's launches, without the host-side launch gaps of eager mode.
    achieved = algorithmic FLOPs of those launches (2*B*H*W*64*64*9 each) / time.  peak = the measured
    cuBLAS bf16 BURST figure (these launches are timed alone, for milliseconds); the sustained figure is given beside it."""
    ops = fd.ops
    burst = peaks.get("bf16_tflops")
    sustained = peaks.get("bf16_tflops_sustained")
    which = "measured (MEASURED_PEAKS.json bf16_tflops: burst, the kernel is timed alone for milliseconds)"
    if burst is None:
        burst, sustained, which = 1590.0, 1400.0, "fallback (B200_PROFILING.md dense bf16 burst)"
    calls = []
    orig = ops.conv3x3

    def recording(x, w, **kw):
        calls.append((x, w, kw))
        orig(x, w, **kw)

    ops.conv3x3 = recording
    try:
        eng.run_forward(pl, pl.x)
        eng.run_backward(pl, pl.dy)
        torch.cuda.synchronize()
    finally:
        ops.conv3x3 = orig

    def timed(sel, reps=20):
        def run():
            for x, w, kw in sel:
                orig(x, w, **kw)
        return _time_graph(run, reps) / 1e3          # ms per replay

    flops = lambda c: 2.0 * c[0].shape[0] * c[0].shape[1] * c[0].shape[2] * 64 * 64 * 9
    tot_ms = timed(calls)
    tot_fl = sum(flops(c) for c in calls)
    by_shape = {}
    for key in sorted({(c[0].shape[1], c[0].shape[2]) for c in calls}, reverse=True):
        sel = [c for c in calls if (c[0].shape[1], c[0].shape[2]) == key]
        ms = timed(sel)
        fl = sum(flops(c) for c in sel)
        by_shape[f"{key[0]}x{key[1]}"] = {"launches": len(sel), "avg_us": ms * 1e3 / len(sel),
                                          "tflops": fl / (ms * 1e-3) / 1e12}
    achieved = tot_fl / (tot_ms * 1e-3) / 1e12
    big = by_shape.get("60x60")
    traffic, traffic_src = None, None
    for fname in ("r2_traffic.json", "r1_conv3x3_traffic.json"):
        try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", fname)))
            tj = tj.get("conv3x3_tc_kernel", tj)
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            break
        except Exception:  # noqa: BLE001
            pass
    return {"kernel": "conv3x3_tc_kernel (tcgen05 implicit GEMM; the fwd + dgrad launches of one step outside the "
                      "fused 15x15 chain)",
            "bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
            "peak_source": which, "peak_sustained": sustained,
            "frac_of_sustained": achieved / sustained if sustained else None,
            "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
            "launches": len(calls),
            "avg_launch_us": tot_ms * 1e3 / len(calls), "by_shape": by_shape,
            "largest_shape_frac": (big["tflops"] / burst) if big else None,
            "note": "M=128,N=64,K=16 SS-mode tcgen05.mma is shared-memory-operand bound (6 KB/MMA at 128 B/clk "
                    "= 48 clk vs the 32 clk tensor-pipe floor): ~2/3 of dense peak is the ceiling of a 64-channel conv"}


if __name__ == "__main__":
    main()
