"""Development probe: the BatchNorm streaming kernels on one conv2-sized tensor (N x 16 x 56 x 56 x 144 bf16), for ncu.
    python tools/bn_probe.py [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cstp_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
C, Cp = 144, 144
g = torch.Generator(device="cuda").manual_seed(1)
raw = torch.randn(N, 16, 56, 56, Cp, device="cuda", generator=g).to(torch.bfloat16)
rows = raw.numel() // Cp
st = ops.BNState.alloc(C, Cp, 2, rows // 2, "cuda")
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
act = torch.empty_like(raw)
d = torch.randn(N, 16, 56, 56, Cp, device="cuda", generator=g).to(torch.bfloat16)
gout = torch.empty_like(raw)
dgam, dbet = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("stats+finalize", lambda: ops.bn_forward_stats(raw, st, gamma, beta, rm, rv)),
                 ("apply", lambda: ops.bn_apply(raw, st, act, relu=True)),
                 ("backward", lambda: ops.bn_backward(d, None, raw, st, gamma, dgam, dbet, gout, mask_from_raw=True))):
    fn()
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1) / 3:.4f} ms for {raw.numel() * 2 / 1e9:.2f} GB tensors")
