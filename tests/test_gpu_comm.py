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
        n = 769352                                           # the PoolResnet-medium flat gradient buffer
        ar = par.PeerAllReduce.create(n, dev)
        if ar is None:
            q.put((rank, "unavailable"))
            return
        flat = torch.empty(n, device=dev)
        ok = True
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
