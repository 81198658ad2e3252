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
        self._comm = None
        self._pending = False

    def _reduce(self, piece: torch.Tensor) -> None:
        if self._avg:
            dist.all_reduce(piece, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group)
            piece.mul_(1.0 / self.world)

    def __call__(self, flat: torch.Tensor) -> None:
        if self.world == 1:
            return
        n = flat.numel()
        step = (n + self.chunks - 1) // self.chunks
        for lo in range(0, n, step):
            self._reduce(flat[lo:lo + step])

    def bucket(self, piece: torch.Tensor, events=()) -> None:
        """All-reduce(mean) of one finished part of the gradient buffer, overlapped with the rest of the backward pass:
        on CUDA the collective is queued on a communication stream behind `events` (recorded where the part became final);
        finish() joins it.  Every rank issues the same buckets in the same order (the engine's static backward program)."""
        if self.world == 1 or piece.numel() == 0:
            return
        if not piece.is_cuda:
            self._reduce(piece)
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=piece.device)
        with torch.cuda.stream(self._comm):
            for e in events:
                self._comm.wait_event(e)
            self._reduce(piece)
        self._pending = True

    def finish(self) -> None:
        """The current stream waits for every bucket queued since the last finish()."""
        if self._pending:
            torch.cuda.current_stream().wait_stream(self._comm)
            self._pending = False


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


class BnSyncP2P:
    """BnSync over NVLink peer memory: the same `.world` / `.all_reduce(tensor)` interface, but every exchange is ONE
    single-CTA kernel (`cstp_bn_sync_exchange`, csrc/p2p_sync.cu) that stores the row into every peer's receive buffer
    (torch symmetric memory), publishes a flag and sums the rows in rank order -- a few microseconds instead of an NCCL
    all-reduce's ~25-30, ~140 times per step.  The sum order is the rank order on every rank, so all ranks hold
    bit-identical statistics.  Two channels (buffer set + DEVICE call counter each): one for the stream the step runs on,
    one for the engine's side stream (register_side_stream) -- keyed by role, not by stream handle, so that the step can be
    captured into a CUDA graph on torch's capture stream and replayed; channels are created in call order, which is the
    same on every rank (the first, eager step creates both)."""

    def __init__(self, group=None, row_max: int = 4 * 4096, slots: int = 4):
        if not dist.is_initialized():
            raise RuntimeError("BnSyncP2P needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.row_max, self.slots = int(row_max), int(slots)
        self._chan = {}
        self._side = set()
        self.err = None

    def register_side_stream(self, stream) -> None:
        """Calls issued while `stream` is current use the second channel (the engine's target-network stream)."""
        self._side.add(stream.cuda_stream)

    def _channel(self):
        import torch.distributed._symmetric_memory as symm_mem
        from . import lib as L
        key = 1 if torch.cuda.current_stream().cuda_stream in self._side else 0
        ch = self._chan.get(key)
        if ch is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            nbytes = L.load().cstp_bn_sync_buffer_bytes(self.world, self.slots, self.row_max)
            buf = symm_mem.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, self.group.group_name)
            torch.cuda.synchronize()
            dist.barrier(self.group)             # every buffer is zeroed before any peer writes into it
            if self.err is None:
                self.err = torch.zeros(1, dtype=torch.int32, device=dev)
            ch = dict(buf=buf, hdl=hdl, peers=int(hdl.buffer_ptrs_dev), seq_dev=torch.zeros(1, dtype=torch.int32, device=dev))
            self._chan[key] = ch
        return ch

    def all_reduce(self, t: torch.Tensor) -> None:
        if self.world == 1:
            return
        from . import lib as L
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise L.CstpError("BnSyncP2P exchanges contiguous fp32 CUDA rows")
        ch = self._channel()
        # the call sequence number lives in device memory and is bumped by the kernel (replayable from a CUDA graph)
        L.check(L.load().cstp_bn_sync_exchange(t.data_ptr(), t.numel(), ch["peers"], self.world, self.rank, self.slots,
                                               self.row_max, 0, ch["seq_dev"].data_ptr(), t.data_ptr(), self.err.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))

    def check(self) -> None:
        """Raises if a peer's flag ever timed out (reads one int from the device: call it outside the step)."""
        if self.err is not None and int(self.err.item()) != 0:
            raise RuntimeError(f"SyncBN peer exchange timed out waiting for rank {int(self.err.item()) - 1}")


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
