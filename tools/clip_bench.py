"""Throughput of the GPU clip pipeline at the pretraining batch size, next to the Pillow oracle on the
host cores (one process, as one DataLoader worker would run it).
    python tools/clip_bench.py [B]"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.data_process.clip_plan import PretrainClipSampler  # noqa: E402
from cstp_b200.data_process.gpu_clips import GpuClipPipeline, collate_labels  # noqa: E402
from cstp_b200.synthetic import synthetic_video  # noqa: E402
from oracle.clip_oracle import render_plan  # noqa: E402  (only for the Pillow CPU timing leg below)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 60
W, H, F = 320, 240, 150
random.seed(0)
np.random.seed(0)
torch.manual_seed(0)
host_vids = [synthetic_video(F + 1, W, H, b % 4) for b in range(4)]
vids = [torch.from_numpy(host_vids[b % 4]).cuda() for b in range(B)]
sampler = PretrainClipSampler()
pipe = GpuClipPipeline()
out = None
t_plan, t_gpu = [], []
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(8):
    t0 = time.perf_counter()
    plans = [sampler.plan(F, W, H) for _ in range(B)]
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    a.record()
    out = pipe.assemble(plans, vids, out)
    labels = collate_labels(plans, device="cuda")
    b_.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if it >= 3:
        t_plan.append(t1 - t0)
        t_gpu.append((a.elapsed_time(b_), t2 - t1))
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = pipe.assemble(plans, vids, out)
    torch.cuda.synchronize()
kern_us = sum(e.device_time for e in prof.events() if "clip_assemble" in e.name)
# Pillow oracle on the host: 4 samples
t0 = time.perf_counter()
for i in range(4):
    render_plan(plans[i], host_vids[i % 4])
cpu_per_sample = (time.perf_counter() - t0) / 4
res = {"B": B, "frame": [W, H], "plan_ms_per_batch": 1e3 * float(np.mean(t_plan)),
       "kernel_ms_per_batch": kern_us / 1e3, "in_bytes_touched_per_batch": int(sum(
           (v.box[2] - v.box[0]) * (v.box[3] - v.box[1]) * 3 * 16 for p in plans for v in p.views)),
       "gpu_ms_per_batch_events": float(np.mean([g[0] for g in t_gpu])),
       "host_wall_ms_per_batch": 1e3 * float(np.mean([g[1] for g in t_gpu])),
       "samples_per_s_gpu": B / (float(np.mean([g[1] for g in t_gpu])) + float(np.mean(t_plan))),
       "pillow_cpu_ms_per_sample_1core": 1e3 * cpu_per_sample, "samples_per_s_pillow_1core": 1.0 / cpu_per_sample,
       "out_bytes_per_batch": 2 * B * 3 * 16 * 112 * 112 * 4}
print(json.dumps(res))
