"""Generates tests/golden/*.pt from the UNMODIFIED reference (imported from /root/reference, never copied).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Protocol = SURVEY.md A.2: torch.manual_seed(1) weights, Generator(0) clips/labels, loss_weight 0.1 1 1 1 1,
SGD lr 0.03 momentum 0.9 wd 5e-4, clip_grad_norm_ 18, two steps.  Stored per batch size:
  * scalars: losses, grad-norm, parameter checksums before/after, label vectors
  * per hooked module call: (mean, std, l2) + 512 sampled values of the output and of its gradient
  * per trainable parameter: gradient l2 norm + 256 sampled values; post-step sampled values
`step_struct_b4.pt` repeats the protocol on the video-like clips of oracle.structured_batch (B=4, with layers): the
fixture the bf16 per-layer comparison uses (i.i.d. noise clips make every small-batch BatchNorm ill-conditioned).
and NT-Xent anchors (loss + gradient samples) from loss/NTXent.py for rows in {64, 256, 1024}.
"""
from __future__ import annotations

import os
import sys
import time

import torch
import torch.nn as nn

REF = os.environ.get("CSTP_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.cstp_oracle import structured_batch, synthetic_batch  # noqa: E402

N_ACT_SAMPLES = 512
N_GRAD_SAMPLES = 256


def sample_idx(numel: int, k: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1234 + numel % 9973)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def summarize(t: torch.Tensor, k: int) -> dict:
    f = t.detach().float().reshape(-1)
    return dict(shape=tuple(t.shape), mean=f.mean().item(), std=f.std().item() if f.numel() > 1 else 0.0,
                l2=f.norm().item(), samples=f[sample_idx(f.numel(), k)].clone())


class disk_offload(torch.autograd.graph.saved_tensors_hooks):
    """Parks every large tensor autograd saves for backward in a scratch file, so the UNMODIFIED reference can run its
    batch-60 step (BASELINE config 3: ~55 GB of saved fp32 activations) on a 62 GB host.  Values are untouched."""

    def __init__(self, root: str, min_numel: int = 1 << 20):
        import numpy as np
        os.makedirs(root, exist_ok=True)
        count = [0]

        def pack(t):
            if t.numel() < min_numel or t.dtype != torch.float32 or not t.is_contiguous():
                return t
            count[0] += 1
            path = os.path.join(root, f"t{count[0]}.bin")
            t.detach().numpy().tofile(path)
            return (path, tuple(t.shape))

        def unpack(h):
            if isinstance(h, torch.Tensor):
                return h
            path, shape = h
            t = torch.from_numpy(np.fromfile(path, dtype=np.float32)).view(shape)
            os.remove(path)
            return t
        super().__init__(pack, unpack)


def run_reference(B: int, steps: int, with_layers: bool, batch_fn=synthetic_batch, offload: str | None = None) -> dict:
    import contextlib
    sys.path.insert(0, REF)
    from models.pace import r21d_byol as ref_mod  # the reference, unmodified

    torch.manual_seed(1)
    model = ref_mod.R21DBYOL(pretrain=True)
    model.train()
    x1, x2, labels = batch_fn(B, 0)
    w = [0.1, 1.0, 1.0, 1.0, 1.0]
    crit = nn.CrossEntropyLoss()
    params = list(model.parameters())
    opt = torch.optim.SGD(params, lr=0.03, momentum=0.9, weight_decay=5e-4)   # main_byol.py:229-232
    out = dict(B=B, labels=[l.clone() for l in labels],
               param_sum_online=sum(p.double().sum().item() for p in model.online_net.parameters()),
               param_sum_target=sum(p.double().sum().item() for p in model.target_net.parameters()),
               param_sum_all=sum(p.double().sum().item() for p in model.parameters()), steps=[])
    for step in range(steps):
        acts, act_grads, calls = {}, {}, {}
        handles = []
        if with_layers and step == 0:
            def mk(name):
                def hook(mod, inp, o):
                    i = calls.get(name, 0)
                    calls[name] = i + 1
                    key = f"{name}#{i}"
                    acts[key] = summarize(o, N_ACT_SAMPLES)
                    if o.requires_grad:
                        o.register_hook(lambda g, key=key: act_grads.__setitem__(key, summarize(g, N_ACT_SAMPLES)))
                return hook
            for name, m in model.named_modules():
                if isinstance(m, (nn.Conv3d, nn.BatchNorm3d, nn.BatchNorm1d, nn.Linear, ref_mod.SpatioTemporalResBlock,
                                  nn.AdaptiveAvgPool3d)):
                    handles.append(m.register_forward_hook(mk(name)))
        t0 = time.time()
        # ---- the step body of main_byol.py:60-91 ----
        with (disk_offload(offload) if offload else contextlib.nullcontext()):
            loss_byol, preds = model(x1, x2, o_type="loss_com")
            loss_byol = loss_byol.mean()
            spa, tem, pb, r1, r2 = labels
            ce = [crit(preds[0], spa), crit(preds[1], tem), crit(preds[2], pb), crit(preds[3], pb), crit(preds[4], r1),
                  crit(preds[5], r2)]
            total = (w[0] * loss_byol + w[1] * ce[0] + w[2] * ce[1] + w[3] * ce[2] + w[3] * ce[3] + w[4] * ce[4]
                     + w[4] * ce[5])
        opt.zero_grad()
        total.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 18)
        rec = dict(loss_byol=loss_byol.item(), ce=[c.item() for c in ce], loss_total=total.item(), grad_norm=float(gnorm),
                   logits=[p.detach().clone() for p in preds], seconds=None)
        if with_layers and step == 0:
            # gradients are stored AFTER clipping (what SGD consumes); the clip coefficient is stored to undo it
            rec["clip_coef"] = min(1.0, 18.0 / (float(gnorm) + 1e-6))
            rec["param_grads"] = {n: summarize(p.grad, N_GRAD_SAMPLES) for n, p in model.named_parameters()
                                  if p.grad is not None}
        opt.step()
        rec["seconds"] = time.time() - t0
        for h in handles:
            h.remove()
        if with_layers and step == 0:
            rec["acts"], rec["act_grads"] = acts, act_grads
            rec["params_after"] = {n: summarize(p, N_GRAD_SAMPLES) for n, p in model.named_parameters()}
            rec["buffers_after"] = {n: summarize(b, N_GRAD_SAMPLES) for n, b in model.named_buffers()
                                    if not n.endswith("num_batches_tracked")}
        rec["param_sum_online"] = sum(p.double().sum().item() for p in model.online_net.parameters())
        rec["param_sum_target"] = sum(p.double().sum().item() for p in model.target_net.parameters())
        out["steps"].append(rec)
        print(f"[ref B={B}] step {step}: byol {rec['loss_byol']:.6f} total {rec['loss_total']:.6f} "
              f"gnorm {rec['grad_norm']:.4f} ({rec['seconds']:.1f}s)", flush=True)
    out["state_dict_keys"] = list(model.state_dict().keys())
    out["param_names"] = [n for n, _ in model.named_parameters()]
    out["param_shapes"] = {n: tuple(p.shape) for n, p in model.named_parameters()}
    return out


def run_ntxent() -> dict:
    sys.path.insert(0, REF)
    from loss.NTXent import NTXentLoss  # the reference, unmodified
    res = {}
    for rows, d, tau in [(64, 32, 0.5), (256, 128, 0.1), (1024, 128, 0.1)]:
        g = torch.Generator().manual_seed(rows)
        z = torch.nn.functional.normalize(torch.randn(rows, d, generator=g), dim=1)
        n = rows // 2
        zis = z[n:].clone().requires_grad_(True)
        zjs = z[:n].clone().requires_grad_(True)
        crit = NTXentLoss("cpu", n, tau, True)
        loss = crit(zis, zjs)
        loss.backward()
        dz = torch.cat([zjs.grad, zis.grad], 0)
        res[rows] = dict(rows=rows, d=d, tau=tau, loss=loss.item(), dz_abs_sum=dz.abs().sum().item(),
                         dz=summarize(dz, 512))
        print(f"[ref ntxent rows={rows}] loss {loss.item():.6f} sum|dz| {dz.abs().sum().item():.4f}", flush=True)
    # unnormalised inputs + dot similarity exercise the other branches
    g = torch.Generator().manual_seed(7)
    z = torch.randn(128, 48, generator=g) * 0.7
    for use_cos in (True, False):
        zis = z[64:].clone().requires_grad_(True)
        zjs = z[:64].clone().requires_grad_(True)
        loss = NTXentLoss("cpu", 64, 0.5, use_cos)(zis, zjs)
        loss.backward()
        dz = torch.cat([zjs.grad, zis.grad], 0)
        res[f"raw_cos{int(use_cos)}"] = dict(rows=128, d=48, tau=0.5, loss=loss.item(), dz=summarize(dz, 512),
                                             dz_abs_sum=dz.abs().sum().item())
    return res


def run_finetune(B: int = 4, offload: str | None = None) -> dict:
    """R21DBYOL(pretrain=False, num_classes=101, cls_bn=True): one training step of main_ft_mp.py:196-214 on B
    video-like clips (SGD lr 0.025 momentum 0.9 wd 1e-3, README.md:68-78), then model.eval() logits of the first clips."""
    import contextlib
    sys.path.insert(0, REF)
    from models.pace import r21d_byol as ref_mod  # the reference, unmodified
    torch.manual_seed(1)
    model = ref_mod.R21DBYOL(pretrain=False, num_classes=101, cls_bn=True)
    model.train()
    x = structured_batch(B, 0)[0]
    labels = torch.randint(0, 101, (B,), generator=torch.Generator().manual_seed(11))
    opt = torch.optim.SGD(model.parameters(), lr=0.025, momentum=0.9, weight_decay=1e-3)
    out = dict(B=B, labels=labels.clone(), state_dict_keys=list(model.state_dict().keys()),
               param_sum=sum(p.double().sum().item() for p in model.parameters()))
    with (disk_offload(offload) if offload else contextlib.nullcontext()):
        logits = model(x, o_type="ft_all")
        loss = nn.CrossEntropyLoss()(logits, labels)
    opt.zero_grad()
    loss.backward()
    out["train"] = dict(loss=loss.item(), logits=logits.detach().clone(),
                        param_grads={n: summarize(p.grad, N_GRAD_SAMPLES) for n, p in model.named_parameters()})
    opt.step()
    out["params_after"] = {n: summarize(p, N_GRAD_SAMPLES) for n, p in model.named_parameters()}
    out["buffers_after"] = {n: summarize(b, N_GRAD_SAMPLES) for n, b in model.named_buffers()
                            if not n.endswith("num_batches_tracked")}
    model.eval()
    with torch.no_grad():
        out["eval_logits_b1"] = model(x[:1], None, o_type="test").clone()
        out["eval_logits_b4"] = model(x[:4], None, o_type="test").clone()
    print(f"[ref finetune] loss {loss.item():.6f} eval argmax {out['eval_logits_b4'].argmax(1).tolist()}", flush=True)
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    which = sys.argv[1:] or ["ntxent", "b2", "b4", "struct_b4", "finetune"]
    if "ntxent" in which:
        torch.save(run_ntxent(), os.path.join(gold, "ntxent_ref.pt"))
    if "b2" in which:
        torch.save(run_reference(2, 2, True), os.path.join(gold, "step_b2.pt"))
    if "b4" in which:
        torch.save(run_reference(4, 2, False), os.path.join(gold, "step_b4.pt"))
    if "struct_b4" in which:       # video-like clips (oracle.structured_batch): the bf16 per-layer parity fixture
        torch.save(run_reference(4, 2, True, structured_batch), os.path.join(gold, "step_struct_b4.pt"))
    if "finetune" in which:
        torch.save(run_finetune(), os.path.join(gold, "finetune_b4.pt"))
    # BASELINE configs 3 / 4 at their full batch (opt-in: minutes of CPU and ~60 GB of scratch disk; one step, with layers)
    scratch = os.environ.get("CSTP_GOLDEN_SCRATCH", "/tmp/cstp_golden_scratch")
    if "b60" in which:
        torch.save(run_reference(60, 1, True, structured_batch, offload=scratch), os.path.join(gold, "step_struct_b60.pt"))
    if "finetune_b60" in which:
        torch.save(run_finetune(60, offload=scratch), os.path.join(gold, "finetune_b60.pt"))
    print("golden fixtures written to", gold)
