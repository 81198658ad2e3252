"""Batch source of the pretraining loop on the GPU clip pipeline: what `get_data_loader` + `UcfRepreBYOLSpPre` /
`Kin400RepreLMDB` + `DistributedSampler` + `data_prefetcher` provide in the reference (utils.py:107-118,
data_process/datasets.py:806-948, 1253-1405, main_byol.py:33-49), with the pixel work on the device.

  * `GpuVideoStore`   decoded videos as uint8 [F][H][W][3] tensors in HBM (JPEG / video decoding stays outside: the
                      reference reads pre-extracted frames, datasets.py:888; feed the store from any decoder).
  * `PretrainBatches` per epoch: the DistributedSampler's index order for this rank, one clip plan per sample drawn in
                      the reference's order, one `cstp_clip_assemble` launch per batch; yields
                      (clip_1, clip_2, (spa, tem, pb, rot_1, rot_2)) -- the argument order of `R21DBYOL.train_step`
                      and of `cstp_b200.train.pretrain_epochs` -- or, with `reference_format=True`, the DataLoader's own
                      `([clip_1, clip_2], [spa, tem, pb, [rot_1, rot_2]])` (main_byol.py:43-49).
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Sequence

import torch

from .clip_plan import PretrainClipSampler
from .gpu_clips import GpuClipPipeline, collate_labels


class GpuVideoStore:
    """index -> CUDA uint8 tensor [F][H][W][3].  `capacity_bytes` bounds the resident set (least recently added videos
    are dropped first); a B200 holds ~2.7 M frames of 171x128 next to the training state."""

    def __init__(self, device="cuda", capacity_bytes: int = 100 << 30):
        self.device = torch.device(device)
        self.capacity = int(capacity_bytes)
        self.bytes = 0
        self._vids: Dict[int, torch.Tensor] = {}

    def put(self, idx: int, frames) -> None:
        t = torch.as_tensor(frames)
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
            raise ValueError("frames must be uint8 [F][H][W][3]")
        t = t.to(self.device, non_blocking=True).contiguous()
        if idx in self._vids:
            self.bytes -= self._vids.pop(idx).numel()
        while self._vids and self.bytes + t.numel() > self.capacity:
            self.bytes -= self._vids.pop(next(iter(self._vids))).numel()
        self._vids[idx] = t
        self.bytes += t.numel()

    def __contains__(self, idx: int) -> bool:
        return idx in self._vids

    def __getitem__(self, idx: int) -> torch.Tensor:
        return self._vids[idx]

    def __len__(self) -> int:
        return len(self._vids)


def distributed_indices(n: int, epoch: int, rank: int = 0, world: int = 1, seed: int = 0, shuffle: bool = True) -> List[int]:
    """torch.utils.data.DistributedSampler's order for this rank (shuffle, drop_last=False): what utils.py:109-110 builds
    and main_byol.py:261 re-seeds with `set_epoch`."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    total = math.ceil(n / world) * world
    pad = total - len(idx)
    if pad > 0:
        idx += (idx * math.ceil(pad / len(idx)))[:pad]
    return idx[rank:total:world]


class PretrainBatches:
    """`batches(epoch)` for `cstp_b200.train.pretrain_epochs`.

    total_frames[i] is the annotated frame count of video i (the third column of trainlist0X_nframe.txt,
    datasets.py:826-827); the store must hold at least that many frames (+ frame 0 unused for the 1-based UCF layout is
    NOT needed: store[i][k] is file number k + 1)."""

    def __init__(self, store: GpuVideoStore, total_frames: Sequence[int], batch_size: int, variant: str = "ucf",
                 sample_duration: int = 16, sample_size: int = 112, rank: int = 0, world: int = 1, seed: int = 0,
                 shuffle: bool = True, drop_last: bool = True, reference_format: bool = False):
        self.store, self.total_frames = store, list(total_frames)
        self.batch_size = int(batch_size)
        self.sampler = PretrainClipSampler(sample_duration, sample_size, variant)
        self.pipe = GpuClipPipeline(sample_duration, sample_size, store.device)
        self.rank, self.world, self.seed, self.shuffle, self.drop_last = rank, world, seed, shuffle, drop_last
        self.reference_format = reference_format
        self._out = [None, None]                  # two output buffer pairs: batch i+1 is assembled while step i runs

    def __call__(self, epoch: int) -> Iterator:
        order = distributed_indices(len(self.total_frames), epoch, self.rank, self.world, self.seed, self.shuffle)
        nb = len(order) // self.batch_size if self.drop_last else math.ceil(len(order) / self.batch_size)
        for b in range(nb):
            ids = order[b * self.batch_size:(b + 1) * self.batch_size]
            plans, vids = [], []
            for i in ids:
                vid = self.store[i]
                plans.append(self.sampler.plan(self.total_frames[i], vid.shape[2], vid.shape[1]))
                vids.append(vid)
            slot = b & 1
            if self._out[slot] is not None and self._out[slot][0].shape[0] != len(ids):
                self._out[slot] = None
            self._out[slot] = self.pipe.assemble(plans, vids, self._out[slot])
            x1, x2 = self._out[slot]
            spa, tem, pb, (r1, r2) = collate_labels(plans, device=self.store.device)
            if self.reference_format:
                yield [x1, x2], [spa, tem, pb, [r1, r2]]
            else:
                yield x1, x2, (spa, tem, pb, r1, r2)
