"""The "existing kernels" bar of SURVEY.md 8(d): the same pretraining step through stock PyTorch operators (cuDNN / cuBLAS /
ATen) on the SAME B200 -- the functional restatement of the reference step (oracle/cstp_oracle.py: F.conv3d, F.batch_norm,
autograd, SGD exactly as main_byol.py runs them) moved to the GPU, in fp32 (TF32 off / on) and under bf16 autocast with
channels_last_3d inputs and cudnn.benchmark.  Not a test (not collected); evidence for profiles/:

    python tests/torch_gpu_bar.py [B] > profiles/r01_torch_gpu_bar.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.engine import trainable_param_specs  # noqa: E402
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from oracle import cstp_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 60
LW = [0.1, 1.0, 1.0, 1.0, 1.0]
torch.backends.cudnn.benchmark = True
torch.cuda.set_per_process_memory_fraction(0.92)


def run(mode: str, batch: int, steps: int = 4):
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    state = {k: v.clone().cuda() for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    trainable = [n for n, _ in trainable_param_specs()]
    x1, x2, labels = O.synthetic_batch(batch, 0)
    x1, x2 = x1.cuda(), x2.cuda()
    labels = tuple(l.cuda() for l in labels)
    tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    if mode == "bf16":
        x1 = x1.contiguous(memory_format=torch.channels_last_3d)
        x2 = x2.contiguous(memory_format=torch.channels_last_3d)
    mom: dict = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    out = None
    for i in range(steps):
        ev[i].record()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
            out = O.pretrain_step(state, trainable, x1, x2, labels, LW, 0.03, mom)
    ev[steps].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    best = min(ms[1:])                       # the first step carries cudnn.benchmark's autotuning
    return {"mode": mode, "batch": batch, "ms_per_step": best, "clips_per_s": batch / best * 1e3, "all_ms": ms,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "loss_total": out["loss_total"]}


res = []
for mode, batch in (("bf16", B), ("tf32", B // 2), ("fp32", B // 2)):
    try:
        torch.cuda.reset_peak_memory_stats()
        res.append(run(mode, batch))
    except torch.cuda.OutOfMemoryError as e:       # report instead of dying: the bar is evidence, not a gate
        res.append({"mode": mode, "batch": batch, "error": "out of memory: " + str(e)[:120]})
        torch.cuda.empty_cache()
    print(json.dumps(res[-1]), flush=True)
