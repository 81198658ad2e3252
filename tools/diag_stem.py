"""Diagnostic: batch-60 step against the reference golden, stem tensors first (development tool)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from oracle import cstp_oracle as O  # noqa: E402
from tests.parity import load_golden, rel, sample_idx  # noqa: E402
from tests import local_parity as LP  # noqa: E402

g = load_golden("step_struct_b60.pt")
B = int(sys.argv[1]) if len(sys.argv) > 1 else g["B"]
x1, x2, labels = O.structured_batch(B, 0)
torch.manual_seed(1)
m = R21DBYOL(pretrain=True)
m.engine_options = {"record": True}
m.cuda()
params = {n: p.detach().float().cpu().clone() for n, p in m.named_parameters()}
losses = m.train_step(x1.cuda(), x2.cuda(), tuple(l.cuda() for l in labels), (0.1, 1, 1, 1, 1), lr=0.03).cpu()
eng = m._engine
s0 = g["steps"][0]
print("byol", losses[7].item(), s0["loss_byol"], "gn", eng.norm_out[0].item(), s0["grad_norm"])
if B == g["B"]:
    for n, s in list(s0["param_grads"].items())[:4]:
        gv = eng.train.view(n, eng.grad).reshape(-1)
        got = gv[sample_idx(gv.numel(), 256).to(gv.device)].float().cpu()
        print("   ", n, round(rel(got, s["samples"]), 4), "l2", float(gv.norm()), s["l2"])
units = eng.units
eng.units = [u for u in units if u["tag"] in ("online.conv1.spatial", "online.conv1.temporal")]
torch.set_num_threads(os.cpu_count())
res = LP.check_conv_units(eng, params, "online")
for k, v in res.items():
    print(k, {a: round(b, 6) for a, b in v.items()})
