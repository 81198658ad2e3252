"""Seeded synthetic inputs (SURVEY.md A.2 / 8(d) protocol): clips, pretext labels and decoded-video stand-ins.

Input generators only -- no restatement of any reference algorithm lives here, so bench.py's native arm, smoke() and the
profiling tools can draw exactly the inputs the oracle-based tests use without importing anything under oracle/."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def synthetic_batch(B: int, seed: int = 0, T: int = 16, S: int = 112):
    """SURVEY.md A.2 / 8(d) protocol: seeded clips in [-1,1) and int64 pretext labels, drawn in a fixed order."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(B, 3, T, S, S, generator=g) * 2 - 1
    x2 = torch.rand(B, 3, T, S, S, generator=g) * 2 - 1
    spa = torch.randint(0, 5, (B,), generator=g)
    tem = torch.randint(0, 5, (B,), generator=g)
    pb = torch.randint(0, 4, (B,), generator=g)
    r1 = torch.randint(0, 4, (B,), generator=g)
    r2 = torch.randint(0, 4, (B,), generator=g)
    return x1, x2, (spa, tem, pb, r1, r2)


def structured_batch(B: int, seed: int = 0, T: int = 16, S: int = 112):
    """Synthetic clips with video-like structure for bf16 parity runs: every sample is its own smooth random field
    (low-resolution noise, trilinearly upsampled, squashed into [-1, 1]) and the second view is a shifted, re-contrasted
    copy plus pixel noise, so features differ strongly between samples the way real clips do.  With the i.i.d. uniform
    noise of `synthetic_batch` all samples produce almost identical features and every BatchNorm over a small batch
    amplifies rounding noise without bound -- fine for fp32 anchors, meaningless for a bf16 comparison.
    Labels are drawn exactly as in `synthetic_batch`."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(B, 3, 4, 7, 7, generator=g) * 1.5 + torch.randn(B, 3, 1, 1, 1, generator=g)
    base = torch.tanh(F.interpolate(low, size=(T, S, S), mode="trilinear", align_corners=False))
    x1 = (0.9 * base + 0.1 * (torch.rand(B, 3, T, S, S, generator=g) * 2 - 1)).clamp(-1, 1)
    shifted = torch.roll(base, shifts=(1, S // 8, -(S // 16)), dims=(2, 3, 4)).flip(4)
    gain = 0.6 + 0.4 * torch.rand(B, 1, 1, 1, 1, generator=g)
    x2 = (gain * shifted + 0.1 * (torch.rand(B, 3, T, S, S, generator=g) * 2 - 1)).clamp(-1, 1)
    spa = torch.randint(0, 5, (B,), generator=g)
    tem = torch.randint(0, 5, (B,), generator=g)
    pb = torch.randint(0, 4, (B,), generator=g)
    r1 = torch.randint(0, 4, (B,), generator=g)
    r2 = torch.randint(0, 4, (B,), generator=g)
    return x1.contiguous(), x2.contiguous(), (spa, tem, pb, r1, r2)


def synthetic_video(n_frames: int, w: int, h: int, seed: int) -> np.ndarray:
    """Deterministic uint8 video [n_frames][h][w][3] (integer arithmetic only: identical on every machine)."""
    f = np.arange(n_frames, dtype=np.int64)[:, None, None]
    y = np.arange(h, dtype=np.int64)[None, :, None]
    x = np.arange(w, dtype=np.int64)[None, None, :]
    chans = []
    for c in range(3):
        smooth = (x * (2 + c) + y * (3 - c) + f * (5 + 2 * c) + seed * 17) % 512
        smooth = np.where(smooth > 255, 511 - smooth, smooth)              # triangle wave: no hard wrap edges
        checker = ((x // 8 + y // 8 + f) % 2) * 24
        texture = ((x * y + f * 3 + c) % 7) * 3
        chans.append(np.clip(smooth // 2 + 40 + checker + texture, 0, 255))
    return np.stack(chans, -1).astype(np.uint8)
