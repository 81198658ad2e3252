"""Bring-up harness for the full step engine on a GPU box: golden-scalar parity at B=2/4 and timing at larger B.
Development tool (graded parity tests live in tests/).   python tools/gpu_step_check.py [B_timing ...]
"""
from __future__ import annotations

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from cstp_b200 import ops  # noqa: E402
from cstp_b200.synthetic import synthetic_batch  # noqa: E402


def golden_check(B):
    g = torch.load(os.path.join(ROOT, "tests", "golden", f"step_b{B}.pt"), weights_only=False)
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True).cuda()
    x1, x2, labels = synthetic_batch(B, 0)
    x1, x2 = x1.cuda(), x2.cuda()
    labels = tuple(l.cuda() for l in labels)
    out = {}
    for step in range(2):
        losses = m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
        torch.cuda.synchronize()
        l = losses.tolist()
        ref = g["steps"][step]
        tot = 0.1 * l[7] + l[6]
        out[f"step{step}"] = dict(byol=(l[7], ref["loss_byol"]), ce=list(zip(l[:6], ref["ce"])), total=(tot, ref["loss_total"]),
                                  gnorm=(m._engine.norm_out[0].item(), ref["grad_norm"]))
        if step == 0 and "param_grads" in ref:
            eng = m._engine
            coef = ref["clip_coef"]
            worst = []
            from oracle.make_golden import sample_idx
            for n, s in ref["param_grads"].items():
                gv = eng.train.view(n, eng.grad).reshape(-1)
                idx = sample_idx(gv.numel(), 256).cuda()
                got = gv[idx].cpu() * coef
                rel = ((got - s["samples"]).norm() / s["samples"].norm().clamp_min(1e-30)).item()
                l2 = gv.norm().item() * coef
                worst.append((round(rel, 4), n, round(l2 / max(s["l2"], 1e-30), 4)))
            worst.sort(reverse=True)
            out["worst_param_grads"] = worst[:12]
            out["median_param_grad_rel"] = sorted(w[0] for w in worst)[len(worst) // 2]
    return out


def timing(B, steps=5):
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True).cuda()
    x1, x2, labels = synthetic_batch(B, 0)
    x1, x2 = x1.cuda(), x2.cuda()
    labels = tuple(l.cuda() for l in labels)
    for _ in range(2):
        m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
    torch.cuda.synchronize()
    eng = m._engine
    res = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    l0 = ops.launch_count()
    ev[0].record()
    for _ in range(steps):
        m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
    ev[1].record()
    torch.cuda.synchronize()
    res["ms_per_step"] = ev[0].elapsed_time(ev[1]) / steps
    res["launches_per_step"] = (ops.launch_count() - l0) / steps
    res["samples_per_s"] = B / res["ms_per_step"] * 1e3
    res["tflops"] = B * 339.5e9 / res["ms_per_step"] / 1e9
    # phase breakdown
    def t(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    res["phase_ms"] = dict(
        im2col=t(lambda: eng.load_clips(x1, x2)),
        fwd_online=t(lambda: [op() for op in eng.fwd_online]),
        ema_pack=t(eng.ema),
        fwd_target=t(lambda: [op() for op in eng.fwd_target]),
        heads=t(lambda: [op() for op in eng.fwd_heads]),
        backward=t(eng.backward),
        optimizer=t(lambda: eng.optimizer_step(0.0)),
    )
    res["mem_gb"] = torch.cuda.max_memory_allocated() / 2**30
    return res


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]] or [8]
    for B in (2, 4):
        t0 = time.time()
        try:
            print(f"GOLDEN B={B} ({time.time() - t0:.1f}s): " + json.dumps(golden_check(B)), flush=True)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(f"GOLDEN B={B} FAILED: {e}", flush=True)
    for B in args:
        try:
            print(f"TIMING B={B}: " + json.dumps(timing(B)), flush=True)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(f"TIMING B={B} FAILED: {e}", flush=True)
        torch.cuda.empty_cache()
