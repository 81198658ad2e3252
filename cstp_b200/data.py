"""Host -> device input staging for the pretraining step.

The reference moves every batch with `.cuda(non_blocking=True)` on the compute stream (main_byol.py:52-58) and, in the
finetune driver, overlaps the copy of the NEXT batch on a side stream (`data_prefetcher`, main_ft_mp.py:313-352).  This
is the same idea for the two-clip pretraining batch: pinned host clips + int64 labels are copied into one of two device
buffers on a copy stream while the previous step computes; `next()` hands the compute stream a ready batch.
"""
from __future__ import annotations

import torch


class ClipPrefetcher:
    def __init__(self, batch_iter, device=None):
        """`batch_iter` yields (x1, x2, labels) with pinned host tensors (fp32 clips, int64 label vectors)."""
        self.it = iter(batch_iter)
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.copy = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.turn = 0
        self.pending = None
        self.h2d_bytes = 0
        self._issue()

    def _issue(self):
        try:
            x1, x2, labels = next(self.it)
        except StopIteration:
            self.pending = None
            return
        b = self.turn
        if self.bufs[b] is None:
            self.bufs[b] = (torch.empty_like(x1, device=self.device), torch.empty_like(x2, device=self.device),
                            tuple(torch.empty_like(l, device=self.device) for l in labels))
        d1, d2, dl = self.bufs[b]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self.free[b])          # the step that last read this buffer has finished
            d1.copy_(x1, non_blocking=True)
            d2.copy_(x2, non_blocking=True)
            for d_, h_ in zip(dl, labels):
                d_.copy_(h_, non_blocking=True)
            self.ready[b].record(self.copy)
        self.h2d_bytes = x1.numel() * x1.element_size() + x2.numel() * x2.element_size() + \
            sum(l.numel() * l.element_size() for l in labels)
        self.pending = b

    def next(self):
        """Returns (x1, x2, labels) on the device, ordered after their copy on the current stream, or None at the end.
        Call `done()` after launching the step that consumes them."""
        if self.pending is None:
            return None
        b = self.pending
        torch.cuda.current_stream().wait_event(self.ready[b])
        self._last = b
        self.turn ^= 1
        self._issue()                                    # start copying the next batch while this one computes
        return self.bufs[b]

    def done(self):
        self.free[self._last].record(torch.cuda.current_stream())


class LossReadback:
    """Device -> host read-back of every step's loss vector without stalling the launch thread.

    The reference reads six `.item()`s per step (main_byol.py:77-84), i.e. it synchronises host and device once per step.
    Here the copy of step i's vector into one of `depth` pinned host slots is queued on the compute stream right behind
    the step (so it sees that step's values), and the host only waits for a slot when it comes round again -- `depth`
    steps later, by which time the copy has long finished -- or in `drain()`.  Every step's vector reaches the host, in
    order; the launch thread runs up to `depth` steps ahead of the GPU, as it does without any read-back."""

    def __init__(self, n: int = 8, depth: int = 4):
        self.slots = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.events = [None] * depth
        self.turn = 0
        self.values: list[list[float]] = []
        self.d2h_bytes = 4 * n

    def _collect(self, k: int) -> None:
        if self.events[k] is not None:
            self.events[k].synchronize()
            self.values.append(self.slots[k].tolist())
            self.events[k] = None

    def push(self, dev_vec: torch.Tensor) -> None:
        k = self.turn
        self._collect(k)                                  # the slot's previous occupant (depth steps ago) is consumed first
        self.slots[k].copy_(dev_vec, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        self.turn = (k + 1) % len(self.slots)

    def drain(self) -> list[list[float]]:
        """Waits for every outstanding copy and returns all vectors pushed so far, oldest first."""
        n = len(self.slots)
        for j in range(n):
            self._collect((self.turn + j) % n)
        return self.values
