"""Data parallelism for the train step: one process per GPU, batch sharded across ranks, weights
replicated, ONE all-reduce (sum) of the flat fp32 gradient buffer per step.

The reference has no distributed code (Trainer(gpus=1), train_model.py:47-53).  Its loss is a SUM over
the batch (models/ModelMeta.py:173-176,215), so the data-parallel gradient is the plain sum of the shard
gradients -- no 1/world_size scaling.  Inference needs no collective.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous shard [begin, end) of a batch of n_items for this rank (ragged tail spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_flat(flat: torch.Tensor, src: int = 0):
    """Replicate the flat parameter buffer of rank `src` (start of training)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)


def allreduce_grads(gflat: torch.Tensor):
    """Sum the flat gradient buffer over ranks, in place (one collective per step)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(gflat, op=dist.ReduceOp.SUM)
    return gflat


def max_over_ranks(value: float, device=None) -> float:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value
