"""Per-layer CUDA-event timing of the tensor-core launches of one step (development tool).
    python tools/layer_profile.py [B]   ->  kind, layer, ms, TFLOP/s"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from cstp_b200.synthetic import synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 60
torch.manual_seed(1)
m = R21DBYOL(pretrain=True).cuda()
x1, x2, labels = synthetic_batch(B, 0)
x1, x2 = x1.cuda(), x2.cuda()
labels = tuple(l.cuda() for l in labels)
for _ in range(3):
    m.train_step(x1, x2, labels, (0.1, 1, 1, 1, 1), lr=0.03)
torch.cuda.synchronize()
eng = m._engine
eng.profile_tensor_launches(x1, x2)
tot = {}
for kind, tag, ms, fl, n, kern in eng.last_profile:
    print(f"{kind:11s} {tag:45s} {ms:8.3f} ms {fl / ms / 1e9:8.1f} TF/s  x{n} {kern}")
    k = tot.setdefault(kind, [0.0, 0.0])
    k[0] += ms
    k[1] += fl
for k, (ms, fl) in tot.items():
    print(f"TOTAL {k:11s} {ms:8.3f} ms {fl / ms / 1e9:8.1f} TF/s")
