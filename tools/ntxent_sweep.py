"""BASELINE config 2: NTXentLoss standalone forward+backward sweep, rows = 2*batch in {256..8192}, d = 128, tau = 0.1,
on one B200 through the drop-in module (cstp_b200.loss.NTXent -> cstp_ntxent).  Prints one JSON line per size.
    python tools/ntxent_sweep.py > gpurun_out/ntxent_sweep.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.loss.NTXent import NTXentLoss  # noqa: E402
from cstp_b200 import ops  # noqa: E402

ANCHOR = {256: 5.915076, 1024: 7.327291, 2048: 8.046678, 4096: 8.733852}      # SURVEY.md A.3 (reference outputs)
d, tau = 128, 0.1
for rows in (256, 512, 1024, 2048, 4096, 8192):
    z = torch.nn.functional.normalize(torch.randn(rows, d, generator=torch.Generator().manual_seed(rows)), dim=1).cuda()
    n = rows // 2
    crit = NTXentLoss("cuda", n, tau, True)
    zis, zjs = z[n:].clone().requires_grad_(True), z[:n].clone().requires_grad_(True)

    def step():
        zis.grad = zjs.grad = None
        loss = crit(zis, zjs)
        loss.backward()
        return loss
    for _ in range(5):
        loss = step()
    torch.cuda.synchronize()
    # kernel-only timing through the C ABI (no autograd / cat overhead)
    zc = torch.cat([zjs, zis]).detach().float().contiguous()
    lo, dz = torch.zeros(1, device="cuda"), torch.empty_like(zc)
    ws = torch.empty(ops.ntxent_workspace_floats(rows, d), device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        ops.ntxent(zc, tau, True, lo, dz, ws)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / iters
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    m_ms = e0.elapsed_time(e1) / iters
    flops = 6.0 * rows * rows * d          # fwd 2*rows^2*d + bwd 4*rows^2*d (S recomputed + W.Z), SURVEY.md 8(d)
    # the SIMT fp32 path of the same entry point (small workspace) for comparison
    ws0 = torch.empty(3 * rows + rows * d, device="cuda")
    ops.ntxent(zc, tau, True, lo, dz, ws0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ops.ntxent(zc, tau, True, lo, dz, ws0)
    e1.record()
    torch.cuda.synchronize()
    simt_ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"rows": rows, "d": d, "tau": tau, "loss": loss.item(), "anchor": ANCHOR.get(rows),
                      "kernel_ms_fwd_bwd": k_ms, "module_ms_fwd_bwd": m_ms, "tflops": flops / k_ms / 1e9,
                      "simt_fp32_ms_fwd_bwd": simt_ms, "pairs_per_s": n / (k_ms * 1e-3)}), flush=True)
