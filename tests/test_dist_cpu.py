"""world_size-2 `gloo` test of the data-parallel host logic (no GPU): batch sharding, parameter
broadcast, the single flat-gradient all-reduce (a SUM, because the reference loss is a sum over the
batch, models/ModelMeta.py:173-176,215), and max-over-ranks timing."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "pytorch-face-detection-from-scratch_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    par = importlib.import_module(PKG + ".parallel")
    r, w, _ = par.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # every rank holds a different "shard gradient" of the same 769349-element flat layout
    eng = importlib.import_module(PKG + ".engine").BackboneEngine(64, 3, 480, 480, 10, 10, 8, 2, 6, 0, lambda h: h > 20)
    g = torch.full((eng.n_flat,), float(rank + 1))
    par.allreduce_grads(g)
    ok_sum = bool(torch.all(g == sum(range(1, world + 1))))           # plain sum, no 1/world scaling
    p = torch.arange(eng.n_flat, dtype=torch.float32) * (1 if rank == 0 else -1)
    par.broadcast_flat(p, src=0)
    ok_bcast = bool(torch.equal(p, torch.arange(eng.n_flat, dtype=torch.float32)))
    t = par.max_over_ranks(10.0 + rank)
    b, e = par.shard_batch(129, rank, world)
    q.put((rank, ok_sum, ok_bcast, t, b, e))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_host_logic_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res)
    assert all(abs(r[3] - 11.0) < 1e-9 for r in res)                  # max over ranks
    assert (res[0][4], res[0][5], res[1][4], res[1][5]) == (0, 65, 65, 129)   # ragged shard covers the batch


def test_shard_batch_covers_everything():
    par = importlib.import_module(PKG + ".parallel")
    for n in (0, 1, 7, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [par.shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
