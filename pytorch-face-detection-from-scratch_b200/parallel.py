"""Data parallelism for the train step: one process per GPU, batch sharded across ranks, weights
replicated, ONE all-reduce (sum) of the flat fp32 gradient buffer per step.

The reference has no distributed code (Trainer(gpus=1), train_model.py:47-53).  Its loss is a SUM over
the batch (models/ModelMeta.py:173-176,215), so the data-parallel gradient is the plain sum of the shard
gradients -- no 1/world_size scaling.  Inference needs no collective.
"""
from __future__ import annotations

import ctypes
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


class PeerAllReduce:
    """Sum-all-reduce of a flat fp32 CUDA buffer through NVLink peer memory: ONE kernel per call
    (csrc/comm.cu), no NCCL and no host work on the data path, so it is captured inside the train step's CUDA graph.

    torch.distributed is the plumbing only: it carries the 64-byte CUDA IPC handles of the per-rank windows at
    start-up.  ``PeerAllReduce.create`` returns None when peer memory cannot be set up on this box (all ranks
    agree on that through one MIN all-reduce); callers then use the NCCL all-reduce of ``allreduce_grads``.
    """

    def __init__(self, numel: int, rank: int, world: int, windows, own):
        self.numel, self.rank, self.world, self._own = numel, rank, world, own
        self.blocks = 64                             # thread blocks of the exchange kernel (SplitAllReduce: 16 for the overlapped half)
        self._windows = windows                      # python ints, entry `rank` is the local window
        self._array = (ctypes.c_void_p * world)(*windows)

    @classmethod
    def create(cls, numel: int, device) -> "PeerAllReduce | None":
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return None
        from .native import lib
        L = lib()
        torch.cuda.set_device(torch.device(device))      # fd_comm_alloc allocates on the CURRENT device
        rank, world = dist.get_rank(), dist.get_world_size()
        ok, own, handle = 1, ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        nbytes = L.fd_comm_window_bytes(numel, world)
        if nbytes <= 0 or L.fd_comm_alloc(nbytes, ctypes.byref(own)) != 0 or L.fd_comm_export(own, handle) != 0:
            ok = 0
        every = [None] * world
        dist.all_gather_object(every, (ok, bytes(handle)))           # plumbing only: 64-byte handles
        windows = [None] * world
        if ok and all(o for o, _ in every):
            for r in range(world):
                if r == rank:
                    windows[r] = own.value
                    continue
                peer = ctypes.c_void_p()
                h = (ctypes.c_ubyte * 64)(*every[r][1])
                if L.fd_comm_import(h, ctypes.byref(peer)) != 0:
                    ok = 0
                    break
                windows[r] = peer.value
        else:
            ok = 0
        torch.cuda.synchronize()
        flags = [None] * world
        dist.all_gather_object(flags, ok)                            # also the "every window is zeroed and mapped" barrier
        if not all(flags):
            for r, w in enumerate(windows):
                if w is not None and r != rank:
                    L.fd_comm_release(ctypes.c_void_p(w))
            if own.value:
                L.fd_comm_free(own)
            return None
        return cls(numel, rank, world, windows, own)

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        from .native import check, cur_stream, dptr, lib
        assert flat.numel() == self.numel and flat.dtype == torch.float32
        check(lib().fd_allreduce_sum_f32_blocks(self._array, self.rank, self.world, dptr(flat, torch.float32), self.numel,
                                                int(self.blocks), cur_stream()), "fd_allreduce_sum_f32")
        return flat

    def status(self) -> int:
        """0 = every exchange completed; 1 + r = a wait on rank r timed out (the peer died or never launched)."""
        from .native import check, lib
        st = ctypes.c_int(0)
        check(lib().fd_comm_status(self._own, ctypes.byref(st)), "fd_comm_status")
        return int(st.value)

    def close(self):
        from .native import lib
        L = lib()
        torch.cuda.synchronize()
        bad = self.status()
        if bad:
            raise RuntimeError(f"peer all-reduce: rank {self.rank} timed out waiting for rank {bad - 1}; "
                               "gradients of at least one step were not reduced")
        if dist.is_initialized():
            dist.barrier()                                   # nobody unmaps a window a peer may still write
        for r, w in enumerate(self._windows):
            if r != self.rank and w:
                L.fd_comm_release(ctypes.c_void_p(w))
        L.fd_comm_free(self._own)
        self._windows = []


class SplitAllReduce:
    """The data-parallel gradient sum of a BackboneEngine step as TWO exchanges over the engine's gradient allocation
    (``engine.exchange_regions``): ``early`` -- the packed weight-gradient accumulators of the fused 15x15 chain, final
    ~2/3 of the backward pass before its end, all-reduced on a side stream while the wide layers are still being
    differentiated -- and ``late`` -- everything else, just before the accumulators are unpacked.  Each exchange is one
    ``PeerAllReduce`` kernel with its own peer window (the two may be in flight at the same time); with
    ``use_nccl=True`` / when peer memory cannot be mapped they are ``dist.all_reduce`` calls (not graph-capturable)."""

    def __init__(self, early_ar, late_ar):
        self._early, self._late = early_ar, late_ar
        self.peer = isinstance(late_ar, PeerAllReduce)

    @classmethod
    def create(cls, engine, plan, device, use_nccl: bool = False) -> "SplitAllReduce | None":
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return None
        early_t, late_t = engine.exchange_regions(plan)
        if not use_nccl:
            late = PeerAllReduce.create(late_t.numel(), device)
            early = PeerAllReduce.create(early_t.numel(), device) if (late is not None and early_t.numel()) else None
            if late is not None and (early is not None or early_t.numel() == 0):
                if early is not None:
                    early.blocks = 16            # runs beside the convolution kernels: leave them the other 132 SMs
                return cls(early, late)
            if late is not None:
                late.close()
        return cls(allreduce_grads, allreduce_grads)

    def early(self, t: torch.Tensor):
        if t.numel():
            self._early(t)

    def late(self, t: torch.Tensor):
        self._late(t)

    def status(self) -> int:
        if not self.peer:
            return 0
        return self._late.status() or (self._early.status() if self._early is not None else 0)

    def close(self):
        if self.peer:
            if self._early is not None:
                self._early.close()
            self._late.close()


def max_over_ranks(value: float, device=None) -> float:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value
