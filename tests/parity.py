"""Helpers shared by the parity tests: layout conversion between the engine's padded NDHWC tensors and the
reference's NCDHW tensors, relative-error metric, golden-fixture sampling."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def to_ncdhw(t: torch.Tensor, C: int) -> torch.Tensor:
    """engine (N, T, H, W, Cp) or (rows, Cp) -> reference (N, C, T, H, W) or (rows, C), fp32 on CPU."""
    t = t.detach().float().cpu()
    if t.dim() == 5:
        return t[..., :C].permute(0, 4, 1, 2, 3).contiguous()
    return t[..., :C].contiguous()


def sample_idx(numel: int, k: int) -> torch.Tensor:
    """Same index draw as oracle/make_golden.py (kept in sync by tests/test_oracle_golden.py)."""
    g = torch.Generator().manual_seed(1234 + numel % 9973)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def load_golden(name: str):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)
