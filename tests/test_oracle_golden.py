"""Pins the CPU oracle (oracle/cstp_oracle.py) against vectors produced by the UNMODIFIED reference.

The reference ships no tests or golden files, so the pin is the reference itself executed in the
build container by oracle/make_golden.py (tests/golden/*.pt).  These tests run on CPU only and never read
/root/reference.
"""
import os

import pytest
import torch

from oracle import cstp_oracle as O
from oracle import make_golden as MG
from tests.parity import load_golden, rel, sample_idx

LW = [0.1, 1.0, 1.0, 1.0, 1.0]


def _init_state():
    """The drop-in module's constructor must reproduce the reference's seeded initialisation draw for draw."""
    from cstp_b200.engine import trainable_param_specs
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    state = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    return m, state, [n for n, _ in trainable_param_specs()]


def test_sample_index_protocol_in_sync():
    for n in (5, 1000, 123457):
        assert torch.equal(sample_idx(n, 256), MG.sample_idx(n, 256))


def test_module_init_and_names_match_reference():
    g = load_golden("step_b2.pt")
    m, state, trainable = _init_state()
    assert list(m.state_dict().keys()) == g["state_dict_keys"]                      # SURVEY.md A.5 order
    assert [n for n, _ in m.named_parameters()] == g["param_names"]
    assert {n: tuple(p.shape) for n, p in m.named_parameters()} == g["param_shapes"]
    on = sum(p.double().sum().item() for p in m.online_net.parameters())
    tg = sum(p.double().sum().item() for p in m.target_net.parameters())
    assert abs(on - g["param_sum_online"]) < 1e-9 and abs(tg - g["param_sum_target"]) < 1e-9
    # trainable order == the reference's parameters() order with the frozen target_net removed
    assert trainable == [n for n in g["param_names"] if not n.startswith("target_net.")]


def test_labels_bit_exact():
    for name, fn in (("step_b2.pt", O.synthetic_batch), ("step_b4.pt", O.synthetic_batch),
                     ("step_struct_b4.pt", O.structured_batch)):
        g = load_golden(name)
        labels = fn(g["B"], 0)[2]
        assert all(a.dtype == torch.int64 and torch.equal(a, b) for a, b in zip(labels, g["labels"]))


def test_oracle_two_steps_match_reference_b2():
    """Losses, grad-norm, sampled parameter gradients (step 0) and the post-step checksums over two SGD steps."""
    g = load_golden("step_b2.pt")
    _, state, trainable = _init_state()
    x1, x2, labels = O.synthetic_batch(2, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    mom: dict = {}
    for step in range(2):
        r = O.pretrain_step(state, trainable, x1, x2, labels, LW, 0.03, mom)
        ref = g["steps"][step]
        assert abs(r["loss_byol"] - ref["loss_byol"]) < 2e-5 * abs(ref["loss_byol"])
        assert abs(r["loss_total"] - ref["loss_total"]) < 2e-5 * abs(ref["loss_total"])
        for a, b in zip(r["ce"], ref["ce"]):
            assert abs(a - b) < 2e-5 * abs(b)
        assert abs(r["grad_norm"] - ref["grad_norm"]) < 1e-3 * ref["grad_norm"]
        for a, b in zip(r["logits"], ref["logits"]):
            assert torch.equal(a.argmax(1), b.argmax(1))                # integer predictions bit-exact
            assert rel(a, b) < 1e-3
        if "param_grads" in ref:
            coef = ref["clip_coef"]
            errs = {}
            for n, s in ref["param_grads"].items():
                if s["l2"] < 1e-4 * ref["grad_norm"] * coef:
                    continue        # Linear biases in front of a BatchNorm: mathematically zero, fp32 noise only
                gv = r["grads"][n].reshape(-1)
                errs[n] = rel(gv[sample_idx(gv.numel(), 256)] * coef, s["samples"])
            assert max(errs.values()) < 2e-2, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
            assert sorted(errs.values())[len(errs) // 2] < 2e-3
        on = sum(v.double().sum().item() for k, v in state.items()
                 if k.startswith("online_net.") and "running" not in k)
        tg = sum(v.double().sum().item() for k, v in state.items()
                 if k.startswith("target_net.") and "running" not in k)
        assert abs(on - ref["param_sum_online"]) < 2e-2, (on, ref["param_sum_online"])
        assert abs(tg - ref["param_sum_target"]) < 1e-4, (tg, ref["param_sum_target"])   # EMA: exact arithmetic


def test_oracle_layers_match_reference_struct_b4():
    """Per hooked module call of the reference (output samples and output-gradient samples), video-like clips."""
    g = load_golden("step_struct_b4.pt")
    _, state, trainable = _init_state()
    x1, x2, labels = O.structured_batch(4, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    tape = O.Tape(True)
    r = O.pretrain_step(state, trainable, x1, x2, labels, LW, 0.03, {}, tape=tape)
    s0 = g["steps"][0]
    assert abs(r["loss_byol"] - s0["loss_byol"]) < 2e-5 * s0["loss_byol"]
    assert abs(r["grad_norm"] - s0["grad_norm"]) < 1e-3 * s0["grad_norm"]
    checked = 0
    for key, s in s0["acts"].items():
        name, call = key.rsplit("#", 1)
        if not name.startswith("online_net.") or not name.endswith(("spatial_conv", "temporal_conv")):
            continue
        tname = f"online.v{int(call) + 1}." + name[len("online_net."):]
        t = tape.acts[tname].detach().reshape(-1)
        assert tuple(tape.acts[tname].shape) == tuple(s["shape"])
        assert rel(t[sample_idx(t.numel(), 512)], s["samples"]) < 1e-3, key
        if key in s0["act_grads"] and tname in tape.act_grads and tape.act_grads[tname] is not None:
            gt = tape.act_grads[tname].reshape(-1)
            assert rel(gt[sample_idx(gt.numel(), 512)], s0["act_grads"][key]["samples"]) < 2e-2, key
        checked += 1
    assert checked == 48        # 24 convolutions x 2 views
    # post-step parameters and BatchNorm running statistics
    for n, s in s0["params_after"].items():
        v = state[n].detach().reshape(-1)
        assert rel(v[sample_idx(v.numel(), 256)], s["samples"]) < 1e-4, n
    for n, s in s0["buffers_after"].items():
        v = state[n].detach().reshape(-1)
        assert rel(v[sample_idx(v.numel(), 256)], s["samples"]) < 1e-4, n


@pytest.mark.parametrize("key", [64, 256, 1024, "raw_cos1", "raw_cos0"])
def test_ntxent_forms_match_reference(key):
    """loss/NTXent.py outputs (tests/golden/ntxent_ref.pt) vs the literal restatement and the closed form (A.3)."""
    g = load_golden("ntxent_ref.pt")[key]
    rows, d, tau = g["rows"], g["d"], g["tau"]
    if isinstance(key, int):
        gen = torch.Generator().manual_seed(rows)
        z = torch.nn.functional.normalize(torch.randn(rows, d, generator=gen), dim=1)
        use_cos = True
    else:
        z = torch.randn(128, 48, generator=torch.Generator().manual_seed(7)) * 0.7
        use_cos = key.endswith("1")
    n = rows // 2
    zis, zjs = z[n:].clone().requires_grad_(True), z[:n].clone().requires_grad_(True)
    closed = O.ntxent_closed_form(zis, zjs, tau, use_cos)
    assert abs(closed.item() - g["loss"]) < 1e-5 * abs(g["loss"])
    if rows <= 256:
        lit = O.ntxent_reference_form(zis.detach(), zjs.detach(), tau, use_cos)
        assert abs(lit.item() - g["loss"]) < 1e-5 * abs(g["loss"])
    closed.backward()
    dz = torch.cat([zjs.grad, zis.grad], 0).reshape(-1)
    assert rel(dz[sample_idx(dz.numel(), 512)], g["dz"]["samples"]) < 1e-4
    assert abs(dz.abs().sum().item() - g["dz_abs_sum"]) < 1e-4 * g["dz_abs_sum"]
