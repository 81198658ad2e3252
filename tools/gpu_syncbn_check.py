"""World-synchronised BatchNorm on real GPUs (run under torchrun with 2+ ranks): a data-parallel step with
parallel.BnSync on per-rank shards against the single-GPU step over the global batch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_syncbn_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cstp_b200 import parallel  # noqa: E402
from cstp_b200.models.pace.r21d_byol import R21DBYOL  # noqa: E402
from cstp_b200.synthetic import structured_batch  # noqa: E402

rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
GB, T, S = 4 * world, 8, 64
LW = (0.1, 1, 1, 1, 1)
x1, x2, labels = structured_batch(GB, 0, T, S)
lo, hi = parallel.shard_bounds(GB, rank, world)


def run(xa, xb, lab, bn_sync, sync):
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True).cuda()
    if bn_sync is not None:
        m.engine_options = {"bn_sync": bn_sync}
    losses = m.train_step(xa.cuda().contiguous(), xb.cuda().contiguous(), tuple(l.cuda().contiguous() for l in lab), LW,
                          lr=0.03, grad_sync=sync).clone()
    torch.cuda.synchronize()
    return m, losses


P2P = "--p2p" in sys.argv
bsync = parallel.BnSyncP2P() if P2P else parallel.BnSync()
m, losses = run(x1[lo:hi], x2[lo:hi], tuple(l[lo:hi] for l in labels), bsync, parallel.GradSync())
if P2P:
    bsync.check()
mean_losses = losses.clone()
dist.all_reduce(mean_losses)
mean_losses /= world
gn = m._engine.norm_out[0].item()
if rank == 0:
    ref, ref_losses = run(x1, x2, labels, None, None)
    sd, rd = m.state_dict(), ref.state_dict()
    rs = max(((sd[k] - rd[k]).norm() / rd[k].norm().clamp_min(1e-12)).item() for k in sd if "running" in k)
    out = {"world": world, "p2p": P2P, "byol": [mean_losses[7].item(), ref_losses[7].item()],
           "total_ce": [mean_losses[6].item(), ref_losses[6].item()], "grad_norm": [gn, ref._engine.norm_out[0].item()],
           "worst_running_stat_rel": rs}
    print("SYNCBN " + json.dumps(out), flush=True)
    assert abs(out["byol"][0] - out["byol"][1]) < 2e-3 * out["byol"][1]
    assert abs(out["total_ce"][0] - out["total_ce"][1]) < 2e-3 * out["total_ce"][1]
    assert rs < 2e-2
dist.barrier()
if P2P:
    # a few more steps on the same model: slot reuse on both streams, sequence numbers far beyond the slot count
    for _ in range(3):
        more = m.train_step(x1[lo:hi].cuda().contiguous(), x2[lo:hi].cuda().contiguous(),
                            tuple(l[lo:hi].cuda().contiguous() for l in labels), LW, lr=0.03, grad_sync=parallel.GradSync())
    torch.cuda.synchronize()
    bsync.check()
    assert torch.isfinite(more).all()
    if rank == 0:
        print("SYNCBN_P2P_MORE_STEPS ok", more.tolist(), flush=True)
dist.barrier()
dist.destroy_process_group()
