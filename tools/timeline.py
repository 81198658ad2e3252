"""Kernel timeline of one pretraining step (torch.profiler chrome trace -> compact CSV: start_us, dur_us, stream, name)
plus the time during which kernels of both streams were running.  Development tool.
    python tools/timeline.py [B] [out.csv]"""
import json
import os
import re
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from cstp_b200.synthetic import synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 60
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "timeline.csv")
torch.manual_seed(1)
m = R21DBYOL(pretrain=True).cuda()
x1, x2, labels = synthetic_batch(B, 0)
x1, x2 = x1.cuda(), x2.cuda()
labels = tuple(l.cuda() for l in labels)
for _ in range(3):
    m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
    torch.cuda.synchronize()
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "t.json")
    prof.export_chrome_trace(p)
    tr = json.load(open(p))
ev = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
rows = []
for e in ev:
    n = re.sub(r"\(.*$", "", e["name"]).replace("void ", "").replace("cstp::", "")[:48]
    rows.append((e["ts"] - t0, e["dur"], e["args"].get("stream"), n))
with open(out, "w") as f:
    for r in rows:
        f.write("%.1f,%.1f,%s,%s\n" % r)
# overlap accounting
pts = []
for s, d, st, n in rows:
    pts.append((s, 1))
    pts.append((s + d, -1))
pts.sort()
act, last, busy1, busy2 = 0, 0.0, 0.0, 0.0
for t, k in pts:
    if act == 1:
        busy1 += t - last
    elif act >= 2:
        busy2 += t - last
    act += k
    last = t
span = rows[-1][0] + rows[-1][1]
print(f"B={B} kernels={len(rows)} span={span / 1e3:.2f} ms  one-kernel={busy1 / 1e3:.2f} ms  two-or-more={busy2 / 1e3:.2f} ms  "
      f"idle={(span - busy1 - busy2) / 1e3:.2f} ms  sum of durations={sum(r[1] for r in rows) / 1e3:.2f} ms")
