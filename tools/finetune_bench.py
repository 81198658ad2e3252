"""BASELINE config 4: r21d finetune step (batch 60, 101-way head) and batch-1 test-mode inference latency on one B200.
    python tools/finetune_bench.py > gpurun_out/finetune_bench.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from cstp_b200 import ops  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (ops.launch_count() - n0) / iters


torch.manual_seed(1)
out = {}
B = 60
m = R21DBYOL(pretrain=False, num_classes=101, cls_bn=True).cuda().train()
g = torch.Generator().manual_seed(0)
x = (torch.rand(B, 3, 16, 112, 112, generator=g) * 2 - 1).cuda()
y = torch.randint(0, 101, (B,), generator=g).cuda()
ms, launches = timed(lambda: m.finetune_step(x, y, lr=0.025, momentum=0.9, weight_decay=1e-3), 10)
flops = B * (3 * 42.733e9 - 1.224e9)            # 3F - 1.224 GFLOP per clip
out["finetune_step_b60"] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3, "launches_per_step": launches,
                            "step_tflops": flops / ms / 1e9}
m.eval()
for b in (1, 10):
    xb = x[:b].contiguous()
    with torch.no_grad():
        ms, launches = timed(lambda: m(xb, None, o_type="test"), 30, warm=5)
    out[f"test_latency_b{b}"] = {"ms": ms, "clips_per_s": b / ms * 1e3, "launches": launches}
print(json.dumps(out))
