"""Development probe: which parameter gradients differ in bits between the fuse policies after ONE step (the end-to-end
bit-identity claim of tests/test_gpu_config3.py::test_fused_prologue_is_bit_identical_to_two_pass_batchnorm).
    python tools/fuse_bits_diag.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from cstp_b200.synthetic import structured_batch  # noqa: E402

LW = (0.1, 1, 1, 1, 1)
x1, x2, lab = structured_batch(3, 5, 8, 64)
batch = (x1.cuda(), x2.cuda(), tuple(l.cuda() for l in lab))
runs = {}
for tag, opts in (("all", {"fuse_apply": True, "fuse_min_positions": 0, "fuse_policy": "all"}),
                  ("default32", {"fuse_apply": True, "fuse_min_positions": 32 * 32}),
                  ("none", {"fuse_apply": False})):
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    m.engine_options = dict(opts, graph=False)
    m.cuda()
    l = m.train_step(*batch, LW, lr=0.03)
    torch.cuda.synchronize()
    e = m._engine
    runs[tag] = (l.clone(), e.grad.clone(), e.train.slots, e.bufs.data.clone())
ref = runs["none"]
for tag in ("all", "default32"):
    l, g, slots, bufs = runs[tag]
    print(f"== {tag} vs none: losses equal {torch.equal(l, ref[0])}, buffers equal {torch.equal(bufs, ref[3])}, "
          f"grads equal {torch.equal(g, ref[1])}")
    for name, (off, shape) in slots.items():
        n = 1
        for s in shape:
            n *= s
        a, b = g[off:off + n], ref[1][off:off + n]
        if not torch.equal(a, b):
            d = (a - b).abs().max().item()
            print(f"   differs: {name:60s} max|d| {d:.3e}  (max|g| {b.abs().max().item():.3e})")
