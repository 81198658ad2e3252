"""Finetune / test branch on CPU: the oracle against the reference's golden vectors
(tests/golden/finetune_b4.pt, produced by oracle/make_golden.py from the unmodified reference), and the finetune engine
over tests/emulate_ops.py against the oracle (fp32 storage: wiring; train step, eval mode, frozen backbone)."""
import os

import pytest
import torch

from oracle import cstp_oracle as O
from tests.parity import load_golden, rel, sample_idx

LR, MOM, WD = 0.025, 0.9, 1e-3


def _model():
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    torch.manual_seed(1)
    return R21DBYOL(pretrain=False, num_classes=101, cls_bn=True)


def _state(m):
    return {k: v.clone() for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}


def test_oracle_matches_reference_golden():
    g = load_golden("finetune_b4.pt")
    m = _model()
    assert list(m.state_dict().keys()) == g["state_dict_keys"]
    assert abs(sum(p.double().sum().item() for p in m.parameters()) - g["param_sum"]) < 1e-9
    state = _state(m)
    trainable = [n for n, _ in m.named_parameters()]
    x = O.structured_batch(4, 0)[0]
    labels = torch.randint(0, 101, (4,), generator=torch.Generator().manual_seed(11))
    assert torch.equal(labels, g["labels"])
    torch.set_num_threads(os.cpu_count() or 1)
    r = O.finetune_step(state, trainable, x, labels, LR, {}, MOM, WD)
    assert abs(r["loss"] - g["train"]["loss"]) < 1e-5 * g["train"]["loss"]
    assert rel(r["logits"], g["train"]["logits"]) < 1e-4
    errs = {}
    for n, s in g["train"]["param_grads"].items():
        gv = r["grads"][n].reshape(-1)
        if s["l2"] > 1e-6:
            errs[n] = rel(gv[sample_idx(gv.numel(), 256)], s["samples"])
    assert sorted(errs.values())[len(errs) // 2] < 2e-3 and max(errs.values()) < 5e-2, sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    for n, s in g["params_after"].items():
        v = state[n].detach().reshape(-1)
        assert rel(v[sample_idx(v.numel(), 256)], s["samples"]) < 1e-4, n
    for n, s in g["buffers_after"].items():
        v = state[n].detach().reshape(-1)
        assert rel(v[sample_idx(v.numel(), 256)], s["samples"]) < 1e-4, n
    with torch.no_grad():
        ev, _ = O.finetune_forward({k: v.detach() for k, v in state.items()}, x, training=False)
        ev1, _ = O.finetune_forward({k: v.detach() for k, v in state.items()}, x[:1], training=False)
    assert rel(ev, g["eval_logits_b4"]) < 1e-4 and rel(ev1, g["eval_logits_b1"]) < 1e-4
    assert torch.equal(ev.argmax(1), g["eval_logits_b4"].argmax(1))


@pytest.fixture()
def emu():
    from cstp_b200 import engine
    from tests import emulate_ops
    saved = engine.ops, engine.ACT_DTYPE
    engine.ops, engine.ACT_DTYPE = emulate_ops, torch.float32
    try:
        yield engine
    finally:
        engine.ops, engine.ACT_DTYPE = saved


def test_emulated_finetune_engine_matches_oracle(emu):
    B, T, S = 4, 8, 64
    x = O.structured_batch(B, 0, T, S)[0]
    labels = torch.randint(0, 101, (B,), generator=torch.Generator().manual_seed(11))
    m = _model()
    state = _state(m)
    trainable = [n for n, _ in m.named_parameters()]
    # ---- drop-in training step: the unmodified loop body of main_ft_mp.py:196-214
    m.train()
    opt = torch.optim.SGD(m.parameters(), lr=LR, momentum=MOM, weight_decay=WD)
    logits = m(x, o_type="ft_all")
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    opt.zero_grad()
    loss.backward()
    opt.step()
    torch.set_num_threads(os.cpu_count() or 1)
    r = O.finetune_step(state, trainable, x, labels, LR, {}, MOM, WD)
    assert abs(loss.item() - r["loss"]) < 1e-5 * r["loss"]
    assert rel(logits, r["logits"]) < 1e-4
    sd = m.state_dict()
    gerr = {n: rel(m._engine.train.view(n, m._engine.grad), r["grads"][n]) for n in trainable
            if r["grads"][n].norm() > 1e-6}
    assert sorted(gerr.values())[len(gerr) // 2] < 1e-2 and max(gerr.values()) < 5e-2, sorted(gerr.items(), key=lambda kv: -kv[1])[:4]
    for k, v in state.items():
        if "running" in k:
            assert rel(sd[k], v) < 1e-4, k
    assert sd["cls_bn.num_batches_tracked"].item() == 1 and sd["online_net.bn1.num_batches_tracked"].item() == 1
    # ---- eval mode after the step (running statistics; different batch size -> a second engine)
    m.eval()
    with torch.no_grad():
        ev = m(x[:1], None, o_type="test")
        ref_ev, _ = O.finetune_forward({k: v.detach().clone() for k, v in m.state_dict().items()}, x[:1], training=False)
    assert ev.shape == (1, 101) and rel(ev, ref_ev) < 1e-4
    assert sd["cls_bn.num_batches_tracked"].item() == 1          # eval updates nothing


def test_emulated_fused_step_and_frozen_backbone(emu):
    from cstp_b200.models.pace.r21d_byol import get_fine_tuning_parameters
    B, T, S = 2, 4, 32
    x = O.structured_batch(B, 1, T, S)[0]
    labels = torch.tensor([3, 77])
    a, b = _model(), _model()
    a.train()
    b.train()
    la = a.finetune_step(x, labels, lr=LR, momentum=MOM, weight_decay=WD).clone()
    opt = torch.optim.SGD(b.parameters(), lr=LR, momentum=MOM, weight_decay=WD)
    lb = torch.nn.CrossEntropyLoss()(b(x, o_type="ft_all"), labels)
    opt.zero_grad()
    lb.backward()
    opt.step()
    assert abs(la.item() - lb.item()) < 1e-6
    sa, sb = a.state_dict(), b.state_dict()
    assert max(rel(sa[k], sb[k]) for k in sa if sa[k].dtype.is_floating_point) < 1e-5
    # ft_fc: get_fine_tuning_parameters(model, 5) freezes everything but `classify` (r21d_byol.py:10-35)
    c = _model()
    c.train()
    groups = get_fine_tuning_parameters(c, 5)
    assert sum(1 for g in groups if g.get("lr", 1) != 0.0) == 2
    opt = torch.optim.SGD(groups, lr=LR, momentum=MOM, weight_decay=WD)
    w0 = c.online_net.conv1.spatial_conv.weight.detach().clone()
    lc = torch.nn.CrossEntropyLoss()(c(x, o_type="ft_fc"), labels)
    opt.zero_grad()
    lc.backward()
    opt.step()
    assert not c._engine.backbone_grads
    assert torch.equal(c.online_net.conv1.spatial_conv.weight, w0) and c.online_net.conv1.spatial_conv.weight.grad is None
    assert c.classify.weight.grad is not None and c.classify.weight.grad.abs().sum() > 0


def test_emulated_fused_step_leaves_frozen_parameters_untouched(emu):
    """ft_fc through the FUSED step: everything but `classify` is frozen (cls_bn included, r21d_byol.py:10-35); optim.SGD
    never touches a parameter without a gradient, so neither weight decay nor momentum may move it."""
    from cstp_b200.models.pace.r21d_byol import get_fine_tuning_parameters
    B, T, S = 2, 4, 32
    x = O.structured_batch(B, 1, T, S)[0]
    labels = torch.tensor([3, 77])
    a, b = _model(), _model()
    a.train()
    b.train()
    get_fine_tuning_parameters(a, 5)
    groups = get_fine_tuning_parameters(b, 5)
    before = {n: p.detach().clone() for n, p in a.named_parameters()}
    for _ in range(2):
        a.finetune_step(x, labels, lr=LR, momentum=MOM, weight_decay=WD)
    opt = torch.optim.SGD(groups, lr=LR, momentum=MOM, weight_decay=WD)
    for _ in range(2):
        lb = torch.nn.CrossEntropyLoss()(b(x, o_type="ft_fc"), labels)
        opt.zero_grad()
        lb.backward()
        opt.step()
    for n, p in a.named_parameters():
        if n.startswith("classify."):
            assert not torch.equal(p, before[n]), n
        else:
            assert torch.equal(p, before[n]), n               # backbone AND cls_bn: bit-identical
    sa, sb = a.state_dict(), b.state_dict()
    assert max(rel(sa[k], sb[k]) for k in sa if sa[k].dtype.is_floating_point) < 1e-5


def test_plateau_lr_matches_torch_reduce_on_plateau():
    """main_ft_mp.py:152: ReduceLROnPlateau(optimizer, 'min', patience) -- the scalar restatement follows torch's on a
    sequence with improvements, plateaus longer than the patience and sub-threshold improvements."""
    from cstp_b200.train import PlateauLR
    g = torch.Generator().manual_seed(3)
    seq = [2.0, 1.5, 1.5, 1.49999, 1.6, 1.7, 1.4, 1.4, 1.4, 1.4, 1.4, 1.39, 1.5, 1.5, 1.5, 1.5] + \
          (1.3 + 0.2 * torch.rand(40, generator=g)).tolist()
    for patience in (0, 2, 3):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=0.025)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "min", patience=patience)
        mine = PlateauLR(0.025, patience)
        for v in seq:
            ref.step(v)
            assert mine.step(v) == opt.param_groups[0]["lr"], (patience, v)


def test_emulated_finetune_epochs_train_validate_and_keep_the_best_checkpoint(emu, tmp_path):
    """cstp_b200.train.finetune_epochs = the epoch loop of main_ft_mp.py:160-310 on the fused step: log columns, weighted
    means, eval-mode validation that updates nothing, plateau schedule, save_<epoch>_max.pth replacing the previous best."""
    import types
    from cstp_b200 import train as TR
    B, T, S = 2, 4, 32
    data = [(O.structured_batch(B, s, T, S)[0], torch.tensor([3 + s, 77 - s])) for s in range(2)]
    m = _model()
    opts = types.SimpleNamespace(n_epochs=3, learning_rate=LR, momentum=MOM, weight_decay=WD, lr_patience=0, task="ft_all",
                                 highest_val={"name": -1.0})
    seen = []
    rows_t, rows_v = TR.finetune_epochs(m, lambda e: iter(data), lambda e: iter(data[:1]), opts, result_path=str(tmp_path),
                                        log_train=seen.append)
    assert [r["epoch"] for r in rows_t] == [1, 2, 3] == [r["epoch"] for r in rows_v] and seen == rows_t
    assert list(rows_t[0]) == ["epoch", "loss", "acc", "lr"] and list(rows_v[0]) == ["epoch", "loss", "acc"]
    assert rows_t[0]["lr"] == LR and all(0.0 <= r["acc"] <= 1.0 for r in rows_t + rows_v)
    # the first epoch's mean training loss: two steps from the initial weights, restated with the drop-in path
    ref = _model()
    ref.train()
    opt = torch.optim.SGD(ref.parameters(), lr=LR, momentum=MOM, weight_decay=WD)
    tot = 0.0
    for x, y in data:
        loss = torch.nn.CrossEntropyLoss()(ref(x, o_type="ft_all"), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        tot += loss.item() * B
    assert abs(rows_t[0]["loss"] - tot / (2 * B)) < 1e-5 * tot
    # validation ran in eval mode: BatchNorm counters only moved during training (2 steps x 3 epochs)
    assert m.cls_bn.num_batches_tracked.item() == 6
    # exactly one best checkpoint is kept, named after the epoch that set the record, in the reference's format
    files = sorted(p.name for p in tmp_path.iterdir())
    assert len(files) == 1 and files[0] == next(iter(opts.highest_val)) and files[0].startswith("save_") and files[0].endswith("_max.pth")
    ck = torch.load(tmp_path / files[0], weights_only=False)
    assert set(ck) == {"epoch", "arch", "state_dict", "optimizer"} and all(k.startswith("module.") for k in ck["state_dict"])
