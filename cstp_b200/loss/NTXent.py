"""Drop-in `NTXentLoss` (reference: loss/NTXent.py:5-62) on the fused sm_100a kernel.

Same constructor and call signature; the loss is evaluated in the closed form
mean_i[ logsumexp_{j != i}(s_ij / tau) - s_i,pos(i) / tau ], which equals the reference's
CrossEntropy(sum)/2N over [positive | masked negatives] logits without materialising the 2N x 2N x d broadcast.
`pos(i) = (i + N) mod 2N` over `cat(zjs, zis)` -- the integer index map is identical to the reference's +-N diagonals.

With `world_size > 1` (and gather=True) the embeddings of every rank are all-gathered first, so each rank contrasts
its rows against the global batch (north-star extension; `batch_size` is then the GLOBAL batch, which is what the
reference passes: main_byol.py:191-196).
"""
from __future__ import annotations

import torch

from .. import ops


class _NTXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, temperature, use_cosine, ws):
        rows, d = z.shape
        loss = torch.zeros(1, device=z.device, dtype=torch.float32)
        dz = torch.empty_like(z) if z.requires_grad else None
        ops.ntxent(z, temperature, use_cosine, loss, dz, ws)
        ctx.dz = dz
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return (ctx.dz * g if ctx.dz is not None else None), None, None, None


class NTXentLoss(torch.nn.Module):
    def __init__(self, device, batch_size, temperature, use_cosine_similarity, gather=False):
        super().__init__()
        self.batch_size = batch_size
        self.temperature = temperature
        self.device = device
        self.use_cosine_similarity = bool(use_cosine_similarity)
        self.gather = gather

    def positive_index(self) -> torch.Tensor:
        """pos(i) for i in [0, 2N): the +-N diagonals of loss/NTXent.py:50-52."""
        n = self.batch_size
        return (torch.arange(2 * n) + n) % (2 * n)

    def forward(self, zis, zjs):
        if self.gather and torch.distributed.is_available() and torch.distributed.is_initialized():
            from ..parallel import all_gather_with_grad
            zis, zjs = all_gather_with_grad(zis), all_gather_with_grad(zjs)
        if zis.shape[0] != self.batch_size or zjs.shape[0] != self.batch_size:
            raise RuntimeError(f"NTXentLoss was built for batch_size={self.batch_size}, got {zis.shape[0]} rows")
        if not zis.is_cuda:
            raise ops.L.CstpError("cstp_b200 NTXentLoss runs on CUDA tensors only: there is no CPU fallback")
        z = torch.cat([zjs, zis], dim=0).float().contiguous()
        key = (z.shape[0], z.shape[1], z.device)
        if getattr(self, "_ws_key", None) != key:          # scratch of the kernel, kept across calls
            self._ws = torch.empty(ops.ntxent_workspace_floats(z.shape[0], z.shape[1]), device=z.device, dtype=torch.float32)
            self._ws_key = key
        return _NTXentFn.apply(z, float(self.temperature), self.use_cosine_similarity, self._ws)
