"""The NVLink peer-memory gradient all-reduce (csrc/comm.cu) with two ranks.  With >= 2 GPUs each rank owns a
device (the production layout); on a one-GPU box both ranks share cuda:0 -- CUDA IPC maps the peer window all the
same and the two spinning kernels are time-sliced, so the protocol (scatter-push, flag barriers, owner-reduce,
broadcast-push, epochs across repeated calls, graph replay) is exercised either way.  gloo is the plumbing here
(NCCL refuses two ranks on one device).  Result must be the exact rank-ordered fp32 sum on both ranks."""
import importlib
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from tests.gpu_util import require_cuda

pytestmark = pytest.mark.gpu
PKG = "pytorch-face-detection-from-scratch_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(rank, call, n):
    g = torch.Generator().manual_seed(1000 * rank + call)
    return torch.randn(n, generator=g)


def _worker(rank, world, port, ndev, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank % ndev))
        import torch.distributed as dist
        par = importlib.import_module(PKG + ".parallel")
        dev = torch.device("cuda", rank % ndev)
        torch.cuda.set_device(dev)
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
        ok = True
        # the PoolResnet-medium flat gradient buffer (two-shot scatter / reduce / broadcast kernel) and a buffer of the size
        # of the split exchange's "late" region (one-shot kernel: every rank pushes everything, one barrier, local sum)
        for n in (769352, 180224):
          ar = par.PeerAllReduce.create(n, dev)
          if ar is None:
            q.put((rank, "unavailable"))
            return
          flat = torch.empty(n, device=dev)
          for call in range(4):                                # eager calls: epochs advance
              flat.copy_(_data(rank, call, n))
              ar(flat)
              want = _data(0, call, n)
              for r in range(1, world):
                  want = want + _data(r, call, n)
              ok = ok and torch.equal(flat.cpu(), want)
          side = torch.cuda.Stream()                           # the same launch replayed from a CUDA graph
          side.wait_stream(torch.cuda.current_stream())
          graph = torch.cuda.CUDAGraph()
          with torch.cuda.stream(side):
              with torch.cuda.graph(graph, stream=side):
                  ar(flat)
          torch.cuda.current_stream().wait_stream(side)
          for call in range(4, 7):
              flat.copy_(_data(rank, call, n))
              graph.replay()
              want = _data(0, call, n)
              for r in range(1, world):
                  want = want + _data(r, call, n)
              ok = ok and torch.equal(flat.cpu(), want)
          ok = ok and ar.status() == 0
          ar.close()
        dist.destroy_process_group()
        q.put((rank, "ok" if ok else "mismatch"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"error: {e!r}"))


def test_peer_allreduce_two_ranks_exact_sum():
    require_cuda()
    world, ndev = 2, torch.cuda.device_count()
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ndev, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = dict(q.get(timeout=240) for _ in range(world))
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
    if any(v == "unavailable" for v in res.values()):
        pytest.skip("CUDA IPC peer windows cannot be mapped on this box (the NCCL path is used instead)")
    assert res == {0: "ok", 1: "ok"}, res


def _dp_worker(rank, world, port, ndev, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank % ndev))
        import torch.distributed as dist
        fd = importlib.import_module(PKG)
        par = fd.parallel
        dev = torch.device("cuda", rank % ndev)
        torch.cuda.set_device(dev)
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
        torch.manual_seed(3)
        m = fd.models.PoolResnet.PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10).to(dev).eval()
        eng = m.engine
        eng.bind(dict(m.named_parameters()))
        B = 2
        shards = []
        for r in range(world):
            g = torch.Generator().manual_seed(50 + r)
            x = torch.rand(B, 3, 480, 480, generator=g).to(dev)
            gt = (torch.rand(B, 5, 10, 10, generator=g) * (torch.rand(B, 1, 10, 10, generator=g) < 0.2)).to(dev)
            gt[:, 0] = (gt[:, 0] > 0).float()
            shards.append((x, gt))
        local = []
        for (x, gt) in shards:                               # every rank differentiates every shard WITHOUT an exchange
            eng.train_step(x, gt, dropout=False)
            local.append(eng.gflat.clone())
        want = local[0]
        for r in range(1, world):
            want = want + local[r]
        split = par.SplitAllReduce.create(eng, eng.plan(B, True), dev)
        if split is None or not split.peer:
            q.put((rank, "unavailable"))
            return
        early_t, late_t = eng.exchange_regions(eng.plan(B, True))
        layout_ok = early_t.numel() == 16 * 9 * 64 * 64 and late_t.numel() + early_t.numel() + eng.offsets["w3"][1] == eng.gzero.numel()
        x, gt = shards[rank]
        eng.train_step(x, gt, dropout=False, allreduce=split)                # eager, overlapped exchange
        torch.cuda.synchronize()
        e1 = ((eng.gflat - want).norm() / want.norm()).item()
        graph, pl, _ = eng.capture_train_step(x, gt, dropout=False, allreduce=split)
        eng.gzero.fill_(7.0)
        graph.replay()
        graph.replay()
        torch.cuda.synchronize()
        e2 = ((eng.gflat - want).norm() / want.norm()).item()
        ok = layout_ok and e1 <= 1e-5 and e2 <= 1e-5 and split.status() == 0
        split.close()
        dist.destroy_process_group()
        q.put((rank, "ok" if ok else f"mismatch layout={layout_ok} eager={e1:.3g} graph={e2:.3g}"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, f"error: {e!r} {traceback.format_exc()[-600:]}"))


def test_data_parallel_step_with_overlapped_split_allreduce():
    """Two ranks, one shard each: the train step with parallel.SplitAllReduce (the chain's packed weight-gradient
    accumulators exchanged on a side stream while the backward pass continues, the rest just before the unpack) leaves
    the SUM of the shard gradients in the flat gradient buffer of both ranks -- eager and replayed from one CUDA graph."""
    require_cuda()
    world, ndev = 2, torch.cuda.device_count()
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, ndev, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = dict(q.get(timeout=300) for _ in range(world))
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
    if any(v == "unavailable" for v in res.values()):
        pytest.skip("CUDA IPC peer windows cannot be mapped on this box (the NCCL path is used instead)")
    assert res == {0: "ok", 1: "ok"}, res
