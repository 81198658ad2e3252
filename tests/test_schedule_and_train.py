"""LR schedule against sequences produced by the reference's scheduler class (tests/golden/sched_ref.json, generated in the
build container), and -- on a GPU -- the epoch driver with the reference's checkpoint format: save, resume, continue."""
import json
import os
import types

import pytest
import torch

from tests.parity import GOLDEN


def test_scheduler_matches_reference_sequences():
    from cstp_b200.scheduler.cosine_anneal import CosineAnnealingWarmupRestarts, lr_at
    for c in json.load(open(os.path.join(GOLDEN, "sched_ref.json"))):
        a = c["args"]
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=a["max_lr"])
        s = CosineAnnealingWarmupRestarts(opt, **a)
        lrs = [opt.param_groups[0]["lr"]]
        for _ in range(c["n"]):
            opt.step()
            s.step()
            lrs.append(opt.param_groups[0]["lr"])
        assert max(abs(x - y) for x, y in zip(lrs, c["lrs"])) < 1e-12
        closed = [lr_at(k, a["first_cycle_steps"], a["max_lr"], a["min_lr"], a["warmup_steps"], a["cycle_mult"], a["gamma"])
                  for k in range(c["n"] + 1)]
        assert max(abs(x - y) for x, y in zip(closed, c["lrs"])) < 1e-12


def test_first_epoch_runs_at_min_lr():
    from cstp_b200.train import epoch_lr
    assert epoch_lr(1, 300, 0.03) == 1e-5                       # main_byol.py:252-258: epoch 1 trains at min_lr
    assert abs(epoch_lr(151, 300, 0.03) - 0.03) < 1e-12         # end of the linear warm-up
    assert epoch_lr(300, 300, 0.03) < 1e-4


def _opts(n_epochs):
    return types.SimpleNamespace(n_epochs=n_epochs, learning_rate=0.03, momentum=0.9, weight_decay=5e-4, clip_grad_norm=1,
                                 loss_weight=[0.1, 1.0, 1.0, 1.0, 1.0])


@pytest.mark.gpu
def test_checkpoint_save_resume_continues_bit_identically(tmp_path):
    from cstp_b200 import train as TR
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    from oracle.cstp_oracle import structured_batch

    def batches(epoch):
        for s in range(2):
            x1, x2, lab = structured_batch(2, 10 * epoch + s, 8, 64)
            yield x1.cuda(), x2.cuda(), tuple(l.cuda() for l in lab)

    opts = _opts(4)
    torch.manual_seed(1)
    a = R21DBYOL(pretrain=True).cuda()
    rows_a = TR.pretrain_epochs(a, batches, opts, result_path=str(tmp_path), save_every=2)
    assert [r["epoch"] for r in rows_a] == [1, 2, 3, 4] and rows_a[0]["lr"] == 1e-5
    ck = torch.load(tmp_path / "save_2.pth", weights_only=False)
    assert set(ck) == {"epoch", "arch", "state_dict", "optimizer"} and ck["epoch"] == 3
    assert all(k.startswith("module.") for k in ck["state_dict"]) and len(ck["state_dict"]) == 351
    # the optimiser entry loads into the optimiser the reference would build (main_byol.py:229-232,243-244)
    probe = R21DBYOL(pretrain=True)
    torch.optim.SGD(probe.parameters(), lr=0.03, momentum=0.9, weight_decay=5e-4).load_state_dict(ck["optimizer"])
    # resume: a fresh model continues from save_2.pth exactly where the uninterrupted run went
    torch.manual_seed(123)
    b = R21DBYOL(pretrain=True).cuda()
    x1, _, _ = structured_batch(2, 0, 8, 64)
    begin = TR.load_checkpoint(str(tmp_path / "save_2.pth"), b, example_clip=x1.cuda())
    assert begin == 2
    rows_b = TR.pretrain_epochs(b, batches, opts, begin_epoch=begin + 1)
    assert [r["epoch"] for r in rows_b] == [3, 4]
    for ra, rb in zip(rows_a[2:], rows_b):
        assert ra == rb                                         # identical losses, bit for bit
    sa, sb = a.state_dict(), b.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
