"""One NT-Xent forward + backward at a given size through the C ABI (for `ncu --metrics gpu__time_duration.sum`).
    python tools/ntxent_profile.py [rows] [d]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cstp_b200 import ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
z = torch.nn.functional.normalize(torch.randn(rows, d, generator=torch.Generator().manual_seed(rows)), dim=1).cuda()
lo, dz = torch.zeros(1, device="cuda"), torch.empty_like(z)
ws = torch.empty(ops.ntxent_workspace_floats(rows, d), device="cuda")
for _ in range(3):
    ops.ntxent(z, 0.1, True, lo, dz, ws)
torch.cuda.synchronize()
print("loss", lo.item())
