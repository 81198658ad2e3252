"""Host-side orchestration of the step engine (buffer wiring, program order, residual / accumulate flags, flat
parameter layout, EMA-before-target order, per-view BatchNorm groups) validated on CPU: cstp_b200.engine runs on top of
tests/emulate_ops.py (a torch restatement of every C-ABI entry point) and is compared with the oracle.

fp32 storage isolates wiring from rounding; bf16 storage reproduces the rounding points of the CUDA kernels and is what
the GPU parity tests compare against end to end (tests/test_gpu_step.py).
"""
import os

import pytest
import torch

from oracle import cstp_oracle as O
from tests import local_parity as LP
from tests.parity import rel, to_ncdhw

LW = (0.1, 1.0, 1.0, 1.0, 1.0)
B, T, S = 4, 8, 64


def _run(engine_mod, dtype, steps=1):
    engine_mod.ACT_DTYPE = dtype
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    m.engine_options = {"record": True}
    before = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    batch = O.structured_batch(B, 0, T, S)
    losses = None
    for _ in range(steps):
        losses = m.train_step(batch[0], batch[1], batch[2], LW, lr=0.03).clone()
    return m, before, batch, losses


@pytest.fixture(scope="module")
def oracle_step():
    from cstp_b200.engine import trainable_param_specs
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    state = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    trainable = [n for n, _ in trainable_param_specs()]
    batch = O.structured_batch(B, 0, T, S)
    torch.set_num_threads(os.cpu_count() or 1)
    tape = O.Tape(True)
    ref = O.pretrain_step(state, trainable, batch[0], batch[1], batch[2], list(LW), 0.03, {}, tape=tape)
    return dict(ref=ref, tape=tape, state_after=state, trainable=trainable)


@pytest.fixture(scope="module")
def emu_fp32():
    from cstp_b200 import engine
    from tests import emulate_ops
    saved = engine.ops, engine.ACT_DTYPE
    engine.ops = emulate_ops
    try:
        yield _run(engine, torch.float32)
    finally:
        engine.ops, engine.ACT_DTYPE = saved


@pytest.fixture(scope="module")
def emu_bf16():
    from cstp_b200 import engine
    from tests import emulate_ops
    saved = engine.ops, engine.ACT_DTYPE
    engine.ops = emulate_ops
    try:
        yield _run(engine, torch.bfloat16)
    finally:
        engine.ops, engine.ACT_DTYPE = saved


def test_fp32_wiring_matches_oracle(emu_fp32, oracle_step):
    m, before, batch, losses = emu_fp32
    ref, tape = oracle_step["ref"], oracle_step["tape"]
    eng = m._engine
    assert abs(losses[7].item() - ref["loss_byol"]) < 1e-5 * ref["loss_byol"]
    for i in range(6):
        assert abs(losses[i].item() - ref["ce"][i]) < 1e-5 * ref["ce"][i]
    assert abs(eng.norm_out[0].item() - ref["grad_norm"]) < 1e-3 * ref["grad_norm"]
    for st in ("conv2", "conv3", "conv4", "conv5"):
        refv = torch.cat([tape.acts[f"online.v1.{st}.block1.out"], tape.acts[f"online.v2.{st}.block1.out"]], 0)
        assert rel(to_ncdhw(eng.named[f"online.{st}.block1.out"], refv.shape[1]), refv) < 1e-4
    # Gradients: ReLU masks of pre-activations within fp32 noise of zero flip (BatchNorm beta starts at 0, so the
    # pre-activation density at the kink is maximal); a flipped fraction f costs about sqrt(f) relative error.
    errs = {n: rel(eng.train.view(n, eng.grad), ref["grads"][n]) for n in oracle_step["trainable"]
            if ref["grads"][n].norm() > 1e-4 * ref["grad_norm"]}
    assert sorted(errs.values())[len(errs) // 2] < 1e-2, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert max(errs.values()) < 3e-2
    for got, want in zip(eng.logits6, ref["logits"]):
        assert torch.equal(got[:, :5].argmax(1), want.argmax(1))
        assert rel(got[:, :5], want) < 1e-4
    # EMA is exact fp32 arithmetic; BN buffers advance twice (two views)
    sd = m.state_dict()
    for k, v in oracle_step["state_after"].items():
        if k.startswith("target_net.") and "running" not in k:
            assert torch.equal(sd[k], v), k
        elif "running" in k:
            assert rel(sd[k], v) < 1e-4, k
    assert sd["online_net.bn1.num_batches_tracked"].item() == 2
    assert sd["overlap_spa.1.num_batches_tracked"].item() == 1


def _params_used(m, before):
    online = {k: v for k, v in before.items() if not k.startswith("target_net.")}
    target = {k: v.detach().clone() for k, v in m.state_dict().items() if k.startswith("target_net.")}
    return online, target


def test_local_parity_checker_on_fp32_emulation(emu_fp32):
    m, before, _, _ = emu_fp32
    res = LP.check_all(m._engine, *_params_used(m, before))
    w = LP.worst(res)
    assert w[0] < 1e-4, w
    assert len([t for t in res if t.startswith("online.")]) == 24 and len([t for t in res if t.startswith("target.")]) == 24
    assert all(e.get("pad_zero", 0.0) == 0.0 for e in res.values())


def test_local_parity_of_bf16_restatement(emu_bf16):
    """The per-layer error budget of bf16 storage (what the CUDA kernels must also meet): < 1e-2 everywhere."""
    m, before, _, _ = emu_bf16
    res = LP.check_all(m._engine, *_params_used(m, before))
    w = LP.worst(res)
    assert w[0] < 1e-2, w


def test_bf16_drift_against_fp32_oracle_is_bounded(emu_bf16, oracle_step):
    """End-to-end drift of a bf16-storage pipeline against the fp32 oracle (documented in DESIGN.md 'Numerics')."""
    m, _, _, losses = emu_bf16
    ref, tape = oracle_step["ref"], oracle_step["tape"]
    eng = m._engine
    assert abs(losses[7].item() - ref["loss_byol"]) < 2e-3 * ref["loss_byol"]
    drift = {}
    for st in ("conv2", "conv3", "conv4", "conv5"):
        refv = torch.cat([tape.acts[f"online.v1.{st}.block1.out"], tape.acts[f"online.v2.{st}.block1.out"]], 0)
        drift[st] = rel(to_ncdhw(eng.named[f"online.{st}.block1.out"], refv.shape[1]), refv)
    assert drift["conv2"] < 2e-2 and drift["conv5"] < 0.1, drift
    assert abs(eng.norm_out[0].item() - ref["grad_norm"]) < 0.1 * ref["grad_norm"]


def test_sgd_update_is_consistent_with_engine_gradients(emu_fp32):
    """p_after = p - lr * (clip * g + wd * p) on the first step (momentum buffer = gradient)."""
    m, before, _, _ = emu_fp32
    eng = m._engine
    coef = eng.norm_out[1].item()
    sd = m.state_dict()
    for n in ("online_net.conv1.spatial_conv.weight", "online_net.conv5.block1.bn2.bias", "predictor.net.3.weight",
              "rotate_cls.3.bias"):
        g = eng.train.view(n, eng.grad)
        want = before[n] - 0.03 * (coef * g + 5e-4 * before[n])
        assert rel(sd[n], want) < 1e-6, n
