"""`CosineAnnealingWarmupRestarts` with the constructor and behaviour of the reference's scheduler/cosine_anneal.py:6-88
(stepped once per EPOCH by main_byol.py:269; host-side scalar math).

Behaviour that matters for parity (checked against lr sequences produced by the reference class,
tests/golden/sched_ref.json): the first epoch runs at `min_lr` (the constructor overwrites every group's lr), then
`warmup_steps` epochs of linear warm-up towards `max_lr`, cosine decay back to `min_lr` over the rest of the cycle, and at
every restart the peak is multiplied by `gamma` and the cycle length (beyond the warm-up) by `cycle_mult`.

Written from the closed form (`lr_at`) rather than as incremental state, so the same function also feeds the fused
`train_step(lr=...)` path without an optimizer object.
"""
from __future__ import annotations

import math

import torch
from torch.optim.lr_scheduler import LRScheduler


def _locate(step: int, first: int, warmup, mult: float):
    """(cycle index, position inside the cycle, cycle length) of global step `step` (step >= 0)."""
    cycle, length, pos = 0, first, step
    while pos >= length:
        pos -= length
        cycle += 1
        length = int((length - warmup) * mult) + warmup
    return cycle, pos, length


def lr_at(step: int, first_cycle_steps: int, max_lr: float, min_lr: float, warmup_steps=0, cycle_mult: float = 1.0,
          gamma: float = 1.0) -> float:
    """Learning rate in force after `step` user calls of scheduler.step() (step 0: freshly constructed -> min_lr), i.e.
    the rate of epoch `step + 1` in main_byol.py:264-269."""
    if step < 0:
        return min_lr
    cycle, pos, length = _locate(step, first_cycle_steps, warmup_steps, cycle_mult)
    peak = max_lr * gamma ** cycle
    if pos < warmup_steps:
        return (peak - min_lr) * pos / warmup_steps + min_lr
    return min_lr + (peak - min_lr) * (1 + math.cos(math.pi * (pos - warmup_steps) / (length - warmup_steps))) / 2


class CosineAnnealingWarmupRestarts(LRScheduler):
    def __init__(self, optimizer: torch.optim.Optimizer, first_cycle_steps: int, cycle_mult: float = 1., max_lr: float = 0.1,
                 min_lr: float = 0.001, warmup_steps: int = 0, gamma: float = 1., last_epoch: int = -1):
        assert warmup_steps < first_cycle_steps
        self.first_cycle_steps, self.cycle_mult = first_cycle_steps, cycle_mult
        self.base_max_lr, self.min_lr, self.warmup_steps, self.gamma = max_lr, min_lr, warmup_steps, gamma
        self._steps = last_epoch
        super().__init__(optimizer, last_epoch)            # the base class performs the initial step(): _steps -> 0
        for group in self.optimizer.param_groups:          # the schedule starts from min_lr, whatever the optimizer had
            group["lr"] = min_lr
        self.base_lrs = [min_lr for _ in self.optimizer.param_groups]

    @property
    def cycle(self) -> int:
        return _locate(max(self._steps, 0), self.first_cycle_steps, self.warmup_steps, self.cycle_mult)[0]

    def get_lr(self):
        lr = lr_at(self._steps, self.first_cycle_steps, self.base_max_lr, self.min_lr, self.warmup_steps, self.cycle_mult,
                   self.gamma)
        return [lr for _ in self.optimizer.param_groups]

    def step(self, epoch=None):
        self._steps = self._steps + 1 if epoch is None else int(math.floor(epoch))
        self.last_epoch = self._steps
        for group, lr in zip(self.optimizer.param_groups, self.get_lr()):
            group["lr"] = lr
