"""Development probe: one conv layer with / without the operand prologue at a realistic size, for ncu.
    python tools/prologue_probe.py <case> <fused 0|1> [N]      case: c2t (144->64 3x1x1) | c2s (64->144 1x3x3) | c4s (256->576, 14x14)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cstp_b200 import ops  # noqa: E402

case, fused = sys.argv[1], sys.argv[2] == "1"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16
cfg = {"c2t": (16, 56, 144, 64, (3, 1, 1), (1, 0, 0)), "c2s": (16, 56, 64, 144, (1, 3, 3), (0, 1, 1)),
       "c4s": (4, 14, 256, 576, (1, 3, 3), (0, 1, 1)), "c3t": (8, 28, 288, 128, (3, 1, 1), (1, 0, 0))}[case]
T, S, cin, cout, kernel, pad = cfg
geom = ops.ConvGeom(kernel, (1, 1, 1), pad)
Cip, Cop = ops.pad16(cin), ops.pad16(cout)
g = torch.Generator(device="cuda").manual_seed(1)
raw = torch.zeros(N, T, S, S, Cip, device="cuda", dtype=torch.bfloat16)
raw[..., :cin] = torch.randn(N, T, S, S, cin, device="cuda", generator=g).to(torch.bfloat16)
st = ops.BNState.alloc(cin, Cip, 2, N * T * S * S // 2, "cuda")
st.scale.copy_(torch.randn(2 * Cip, device="cuda", generator=g))
st.shift.copy_(torch.randn(2 * Cip, device="cuda", generator=g) * 0.3)
w = torch.randn(cout, cin, *kernel, device="cuda", generator=g) / (cin * geom.taps) ** 0.5
wp = torch.empty(Cop, geom.taps * ops.pad64(Cip), device="cuda", dtype=torch.bfloat16)
ops.pack_weight(w, wp)
out = torch.empty(N, T, S, S, Cop, device="cuda", dtype=torch.bfloat16)
so = ops.BNState.alloc(cout, Cop, 2, N * T * S * S // 2, "cuda")
plan = ops.conv_fwd_plan(raw, wp, out, geom, stats=so, prologue=st if fused else None)
gg = torch.randn(N, T, S, S, Cop, device="cuda", generator=g).to(torch.bfloat16)
part = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)
spec = ops.wgrad_plan(raw, gg, geom, cout, cin, part, prologue=st if fused else None)
dw = torch.empty_like(w)
wtp = torch.empty(Cip, geom.taps * ops.pad64(Cop), device="cuda", dtype=torch.bfloat16)
ops.pack_weight(w, wtp, transpose=True)
dx = torch.empty(N, T, S, S, Cip, device="cuda", dtype=torch.bfloat16)
dplans, _ = ops.conv_dgrad_plans(gg, wtp, dx, geom)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", plan.run), ("wgrad", lambda: spec.run(dw)), ("dgrad", lambda: [p_.run() for p_ in dplans])):
    for _ in range(2):
        fn()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    kn = ops.kernel_name({"fwd": plan, "wgrad": spec, "dgrad": dplans[0]}[name])
    print(case, "fused" if fused else "plain", name, kn, f"{e0.elapsed_time(e1) / 5:.4f} ms")
