"""Data-parallel plumbing of the pretraining step: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The path shards over samples exactly as the reference does (models/model.py:82-103, utils.py:107-118): weights,
optimiser state and the EMA target are replicated, each rank steps its own `batch_size // world_size` samples, and the
only data-path exchange is the gradient all-reduce (mean) of the flat fp32 gradient buffer -- one collective per step
instead of DDP's bucketed hooks -- plus, optionally, the embedding all-gather that gives NT-Xent global negatives.
BatchNorm statistics stay per GPU: the reference's --sync_bn builds a process group of size one.
"""
from __future__ import annotations

import datetime
import os

import torch
import torch.distributed as dist


def wait_for_cuda(max_wait_s: float = 45.0) -> bool:
    """Probes CUDA in a SUBPROCESS until it initialises (a failed cuInit is sticky inside a process).  On the shared GPU
    boxes the driver was once seen refusing to initialise for a moment right after another process had exited; callers
    (bench.py, __graft_entry__.smoke) use this before they touch CUDA themselves."""
    import subprocess
    import sys
    import time
    t0 = time.time()
    while True:
        r = subprocess.run([sys.executable, "-c", "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"],
                           capture_output=True)
        if r.returncode == 0:
            return True
        if time.time() - t0 > max_wait_s:
            return False
        time.sleep(3.0)


def env_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched as a plain process."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        # fail fast instead of holding the GPUs for NCCL's default 10 minutes when ranks disagree on a collective
        kw = {"timeout": datetime.timedelta(seconds=int(os.environ.get("CSTP_DIST_TIMEOUT_S", "180")))}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def per_rank_batch(global_batch: int, world: int) -> int:
    """utils.py:111 -- int(batch_size / world_size) samples per GPU (remainder dropped, drop_last semantics)."""
    return int(global_batch / world)


def shard_bounds(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous-by-rank slice [lo, hi) of a global batch."""
    b = per_rank_batch(global_batch, world)
    return rank * b, (rank + 1) * b


class GradSync:
    """All-reduce(mean) of the engine's flat gradient buffer; call between backward and the optimiser step.

    chunks > 1 splits the buffer so that NCCL can pipeline the pieces on its own stream; every rank ends with the
    bit-identical mean, so replicated weights never drift."""

    def __init__(self, group=None, chunks: int = 1):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.chunks = max(1, chunks)
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"

    def __call__(self, flat: torch.Tensor) -> None:
        if self.world == 1:
            return
        n = flat.numel()
        step = (n + self.chunks - 1) // self.chunks
        for lo in range(0, n, step):
            piece = flat[lo:lo + step]
            if self._avg:
                dist.all_reduce(piece, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group)
                piece.mul_(1.0 / self.world)


class BnSync:
    """World-synchronised BatchNorm statistics (the north star's SyncBN; the reference's own --sync_bn group holds one
    rank).  Passed to the engine as `bn_sync`: every BatchNorm call exchanges ONE small row
    ([groups][2][Cp] fp32 sums) per direction with an all-reduce(SUM) over the data-parallel group; with it a step over
    world x B samples equals the single-GPU step over the global batch."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def all_reduce(self, t: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)


def broadcast_parameters(tensors, src: int = 0, group=None) -> None:
    """DDP's initial parameter broadcast (models/model.py:97-103) over the engine's flat buffers."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src, group=group)


class _AllGather(torch.autograd.Function):
    """all_gather along dim 0.  Every rank evaluates the SAME deterministic loss on the gathered rows, so the sum over
    ranks of d(loss_r)/d(local rows) is world * (this rank's slice of the upstream gradient): the backward needs no
    collective (the factor cancels against the data-parallel gradient mean)."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        ctx.group, ctx.rows = group, x.shape[0]
        out = torch.empty(world * x.shape[0], *x.shape[1:], device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        r, world = dist.get_rank(ctx.group), dist.get_world_size(ctx.group)
        return g[r * ctx.rows:(r + 1) * ctx.rows] * float(world), None


def all_gather_with_grad(x: torch.Tensor, group=None) -> torch.Tensor:
    """Rank-ordered concatenation of `x` from every rank, differentiable w.r.t. the local slice."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return _AllGather.apply(x, group)
