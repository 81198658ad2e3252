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


def test_checkpoint_optimizer_entry_drives_a_real_torch_sgd(emulated_engine, tmp_path):
    """main_byol.py:243-258 on a checkpoint written by save_checkpoint: the reference builds optim.SGD + the cosine
    scheduler from opts, loads ck['optimizer'] (which REPLACES the parameter groups) and steps.  Every SGD hyper-parameter
    therefore has to be in the saved group; the momentum buffers must land on the right parameters."""
    from cstp_b200 import train as TR
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    from cstp_b200.scheduler.cosine_anneal import CosineAnnealingWarmupRestarts
    from oracle.cstp_oracle import structured_batch
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    x1, x2, lab = structured_batch(2, 0, 4, 32)
    m.train_step(x1, x2, lab, (0.1, 1, 1, 1, 1), lr=0.02)
    path = tmp_path / "save_7.pth"
    TR.save_checkpoint(str(path), torch.nn.DataParallel(m), 7, "r21d_byol-1", lr=0.02, initial_lr=0.03)   # wrapped: one prefix
    ck = torch.load(path, weights_only=False)
    assert all(k.startswith("module.") and not k.startswith("module.module.") for k in ck["state_dict"])
    probe = R21DBYOL(pretrain=True)
    probe.load_state_dict({k[len("module."):]: v for k, v in ck["state_dict"].items()})
    opt = torch.optim.SGD(probe.parameters(), lr=0.03, momentum=0.9, weight_decay=5e-4)
    opt.load_state_dict(ck["optimizer"])
    g = opt.param_groups[0]
    assert (g["lr"], g["momentum"], g["dampening"], g["weight_decay"], g["nesterov"]) == (0.02, 0.9, 0, 5e-4, False)
    sched = CosineAnnealingWarmupRestarts(opt, first_cycle_steps=300, max_lr=0.03, min_lr=1e-5, warmup_steps=150, gamma=0.5)
    names = [n for n, _ in probe.named_parameters()]
    i = names.index("online_net.conv2.block1.conv1.spatial_conv.weight")
    p = list(probe.parameters())[i]
    buf = opt.state[p]["momentum_buffer"].clone()                # (SGD updates the buffer in place)
    assert torch.equal(buf, m._engine.train.view(names[i], m._engine.mom).cpu())
    assert all(list(probe.parameters())[j] not in opt.state for j, n in enumerate(names) if n.startswith("target_net."))
    for q in probe.parameters():
        if q.requires_grad:
            q.grad = torch.zeros_like(q)
    before = p.detach().clone()
    opt.step()                                                  # KeyError here before the group carried its hyper-parameters
    sched.step()
    lr_used = 1e-5                                               # the scheduler's __init__ put the group at min_lr
    want = before - lr_used * (0.9 * buf + 5e-4 * before)
    assert torch.allclose(p.detach(), want, rtol=0, atol=1e-8)
    # and back: the engine takes the momentum buffers of an optimizer state the reference wrote
    m2 = R21DBYOL(pretrain=True)
    m2._bind(x1)
    TR.load_optimizer_state_dict(m2, opt.state_dict())
    assert not m2._engine.first_step
    assert torch.equal(m2._engine.train.view(names[i], m2._engine.mom), opt.state[p]["momentum_buffer"])


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
