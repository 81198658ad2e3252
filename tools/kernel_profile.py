"""Kernel-name breakdown of one pretraining step with torch.profiler (CUPTI, kernels running back to back as in the
real step -- unlike the ncu launch list, which serialises and flushes caches).  Development tool.
    python tools/kernel_profile.py [B]"""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from cstp_b200.synthetic import synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 60
torch.manual_seed(1)
m = R21DBYOL(pretrain=True).cuda()
if "--no-overlap" in sys.argv:          # one stream: kernel durations are not inflated by a concurrent kernel
    m.engine_options = {"overlap": False}
x1, x2, labels = synthetic_batch(B, 0)
x1, x2 = x1.cuda(), x2.cuda()
labels = tuple(l.cuda() for l in labels)
for _ in range(3):
    m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"\(.*$", "", ev.name).replace("void ", "").replace("cstp::", "")[:70]
        agg[n][0] += 1
        agg[n][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"B={B}: {sum(v[0] for v in agg.values())} kernels, sum of kernel time {tot / 1e3:.3f} ms")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:70s} {c:5d} {t / 1e3:9.3f} ms {100 * t / tot:5.1f}%")
