"""Diagnostic: stem weight gradient (packed row pairs, wgrad layout 1) against cuDNN fp32 at several sizes."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200 import ops  # noqa: E402
from cstp_b200.ops import STEM_GEOM, STEM_CHANNELS  # noqa: E402
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(0)
for (N, T, H, W) in ((1, 1, 16, 16), (1, 2, 112, 112), (2, 16, 112, 112), (16, 16, 112, 112), (120, 16, 112, 112)):
    x = torch.rand(N, 3, T, H, W, device=dev, generator=gen) * 2 - 1
    P = torch.empty(N, T, H // 2, W // 2, STEM_CHANNELS, device=dev, dtype=torch.bfloat16)
    ops.stem_pack(x, P)
    g = torch.zeros(N, T, H // 2, W // 2, 96, device=dev, dtype=torch.bfloat16)
    g[..., :83] = torch.randn(N, T, H // 2, W // 2, 83, device=dev, generator=gen).to(torch.bfloat16)
    scratch = torch.empty(ops.wgrad_partials_need(tuple(P.shape), tuple(g.shape), STEM_GEOM), device=dev)
    dw = torch.full((83, 3, 1, 7, 7), 9.0, device=dev)
    spec = ops.wgrad_plan(P, g, STEM_GEOM, 83, STEM_CHANNELS, scratch, layout=1)
    spec.run(dw)
    dw2 = torch.full((83, 3, 1, 7, 7), 9.0, device=dev)
    scratch2 = torch.empty(64 * 4 * 64 * 96 * 4, device=dev)
    spec2 = ops.wgrad_plan(P, g, STEM_GEOM, 83, STEM_CHANNELS, scratch2, layout=1, allow_halo=False)
    spec2.run(dw2)
    w = torch.zeros(83, 3, 1, 7, 7, device=dev, requires_grad=True)
    xb = x.to(torch.bfloat16).float()
    gf = g[..., :83].float().permute(0, 4, 1, 2, 3).contiguous()
    ref = torch.zeros(83, 3, 1, 7, 7, device=dev)
    for n0 in range(0, N, 8):
        y = F.conv3d(xb[n0:n0 + 8], w, stride=(1, 2, 2), padding=(0, 3, 3))
        (gw,) = torch.autograd.grad(y, w, gf[n0:n0 + 8])
        ref += gw
    torch.cuda.synchronize()
    r = lambda a: ((a - ref).norm() / ref.norm()).item()
    print((N, T, H, W), "halo", r(dw), "gemm", r(dw2), "splits", spec.plan.splits, spec2.plan.splits,
          "norms", dw.norm().item(), ref.norm().item(), flush=True)
    # per kh slice
    print("   per kh:", [round(((dw[:, :, 0, k] - ref[:, :, 0, k]).norm() / ref[:, :, 0, k].norm()).item(), 4) for k in range(7)])
