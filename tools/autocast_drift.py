"""Three pipelines, one batch, one B200: how far does bf16 STORAGE move a 24-layer R(2+1)D step away from fp32?

  fp32      the functional restatement of the reference step (oracle/cstp_oracle.py: F.conv3d / F.batch_norm / autograd)
            on the GPU, TF32 off -- the yardstick
  autocast  the very same torch code under torch.autocast(bfloat16): stock PyTorch / cuDNN mixed precision, nothing of ours
  engine    cstp_b200 (bf16 activations, fp32 accumulation and statistics)

Reported per pipeline against fp32: relative error of every conv output of the online network (worst / median, and the last
conv of each stage), of the block outputs, of every parameter gradient (median over tensors, cosine of the whole
gradient), and the losses.  Evidence for DESIGN.md section 3 ("the drift is bf16 storage, not kernels"); test
infrastructure (it runs the oracle), not product code.

    python tools/autocast_drift.py [B] > profiles/r02_drift_three_pipelines.json
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
LW = [0.1, 1.0, 1.0, 1.0, 1.0]
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def oracle_run(autocast: bool, batch):
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    state = {k: v.clone().cuda() for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    trainable = [n for n, _ in trainable_param_specs()]
    x1, x2, labels = batch
    tape = O.Tape(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = O.pretrain_step(state, trainable, x1, x2, labels, LW, 0.03, {}, tape=tape)
    acts = {k: v.detach() for k, v in tape.acts.items() if k.startswith("online.") and
            (k.endswith(("spatial_conv", "temporal_conv", "block1.out")))}
    return out, acts


def summarize(acts, ref_acts, grads, ref_grads, losses, ref_losses):
    conv = {k: rel(acts[k], ref_acts[k]) for k in ref_acts if k.endswith("_conv")}
    order = sorted(conv.values())
    res = {"conv_out_worst": order[-1], "conv_out_median": order[len(order) // 2]}
    for st in ("conv2", "conv3", "conv4", "conv5"):
        res[f"{st}_block_out"] = rel(torch.cat([acts[f"online.v1.{st}.block1.out"], acts[f"online.v2.{st}.block1.out"]]),
                                      torch.cat([ref_acts[f"online.v1.{st}.block1.out"], ref_acts[f"online.v2.{st}.block1.out"]]))
    per = sorted(rel(grads[n], ref_grads[n]) for n in ref_grads if ref_grads[n].norm() > 1e-6 * ref_losses["grad_norm"])
    fa = torch.cat([grads[n].reshape(-1).float() for n in ref_grads])
    fb = torch.cat([ref_grads[n].reshape(-1).float() for n in ref_grads])
    res.update(grad_rel_median=per[len(per) // 2], grad_rel_worst=per[-1],
               grad_cosine=torch.nn.functional.cosine_similarity(fa, fb, dim=0).item(),
               grad_norm_rel=abs(losses["grad_norm"] - ref_losses["grad_norm"]) / ref_losses["grad_norm"],
               loss_total_rel=abs(losses["loss_total"] - ref_losses["loss_total"]) / ref_losses["loss_total"],
               loss_byol_rel=abs(losses["loss_byol"] - ref_losses["loss_byol"]) / ref_losses["loss_byol"])
    return res


def main():
    x1, x2, labels = O.structured_batch(B, 0)
    batch = (x1.cuda(), x2.cuda(), tuple(l.cuda() for l in labels))
    ref, ref_acts = oracle_run(False, batch)
    ref_grads = {k: v.detach() for k, v in ref["grads"].items()}
    ref_l = dict(loss_total=ref["loss_total"], loss_byol=ref["loss_byol"], grad_norm=ref["grad_norm"])
    out = {"batch": B, "clips": "structured_batch(B, 0), 3x16x112x112", "yardstick": "oracle on the GPU, fp32, TF32 off"}
    ac, ac_acts = oracle_run(True, batch)
    out["torch_autocast_bf16"] = summarize(ac_acts, ref_acts, {k: v.detach() for k, v in ac["grads"].items()}, ref_grads,
                                           dict(loss_total=ac["loss_total"], loss_byol=ac["loss_byol"], grad_norm=ac["grad_norm"]),
                                           ref_l)
    del ac, ac_acts
    torch.cuda.empty_cache()
    # ---- the engine
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    m.engine_options = {"record": True}
    m.cuda()
    losses = m.train_step(*batch, tuple(LW), lr=0.03).cpu()
    eng = m._engine
    acts = {}
    for k, refv in ref_acts.items():
        view, name = k.split(".")[1], ".".join(k.split(".")[2:])
        if name.endswith("_conv"):
            tag = "online." + name.replace("_conv", "") + ".raw"
        else:
            tag = "online." + name                       # block outputs: "<stage>.block1.out"
        t = eng.named[tag]
        Bv, C = refv.shape[0], refv.shape[1]
        if t.dim() == 5 and t.shape[0] == 1:
            t = t.view(2 * B, refv.shape[2], refv.shape[3], refv.shape[4], t.shape[-1])
        n0 = 0 if view == "v1" else B
        acts[k] = t[n0:n0 + Bv, ..., :C].permute(0, 4, 1, 2, 3).float()
    grads = {n: eng.train.view(n, eng.grad) for n in ref_grads}
    total = LW[0] * losses[7].item() + losses[6].item()
    out["cstp_b200_engine"] = summarize(acts, ref_acts, grads, ref_grads,
                                        dict(loss_total=total, loss_byol=losses[7].item(), grad_norm=eng.norm_out[0].item()), ref_l)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
