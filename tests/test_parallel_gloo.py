"""Data-parallel host logic on CPU: world_size-2 `gloo` process groups (one process per rank, as under torchrun).

Covers cstp_b200.parallel (gradient all-reduce(mean) of the flat buffer, parameter broadcast, differentiable embedding
all-gather, contiguous-by-rank sharding of utils.py:107-118) and one data-parallel pretraining step of the engine over
tests/emulate_ops.py: both ranks must end with bit-identical weights equal to a single-process step that averages the
two shards' gradients (per-GPU BatchNorm statistics: the reference's --sync_bn group holds one rank)."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

LW = (0.1, 1.0, 1.0, 1.0, 1.0)
GLOBAL_B, T, S = 4, 4, 32


def _init(rank, world, path):
    dist.init_process_group("gloo", init_method=f"file://{path}", rank=rank, world_size=world)


def _worker_collectives(rank, world, path, out):
    from cstp_b200 import parallel as P
    _init(rank, world, path)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    P.GradSync(chunks=3)(flat)
    ok = torch.allclose(flat, torch.arange(1000, dtype=torch.float32) * 1.5)
    w = torch.full((7,), float(rank))
    P.broadcast_parameters([w], src=0)
    ok &= bool((w == 0).all())
    x = (torch.arange(6, dtype=torch.float32).view(3, 2) + 10 * rank).requires_grad_(True)
    g = P.all_gather_with_grad(x)
    ok &= g.shape == (6, 2) and torch.equal(g[3 * rank:3 * rank + 3], x.detach())
    # every rank evaluates the same loss on the gathered rows, so the local slice receives world x its share of the
    # upstream gradient: the data-parallel gradient MEAN over ranks then yields the true global-loss gradient
    wgt = torch.arange(12, dtype=torch.float32).view(6, 2)
    (g * wgt).sum().backward()
    ok &= torch.allclose(x.grad, world * wgt[3 * rank:3 * rank + 3])
    ok &= P.shard_bounds(128, rank, world) == (64 * rank, 64 * rank + 64) and P.per_rank_batch(60, 8) == 7
    torch.save(bool(ok), f"{out}.{rank}")
    dist.destroy_process_group()


def _step(batch, rank_slice, grad_sync=None, bn_sync=None, ntxent=None):
    from cstp_b200 import engine
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    from tests import emulate_ops
    engine.ops = emulate_ops
    engine.ACT_DTYPE = torch.float32
    torch.manual_seed(1)
    m = R21DBYOL(pretrain=True)
    m.engine_options = {}
    if bn_sync is not None:
        m.engine_options["bn_sync"] = bn_sync
    if ntxent is not None:
        m.engine_options["ntxent"] = ntxent
    lo, hi = rank_slice
    x1, x2, labels = batch
    m.train_step(x1[lo:hi].contiguous(), x2[lo:hi].contiguous(), tuple(l[lo:hi].contiguous() for l in labels), LW, lr=0.03,
                 grad_sync=grad_sync)
    return m


def _worker_step(rank, world, path, out):
    from cstp_b200 import parallel as P
    from oracle import cstp_oracle as O
    torch.set_num_threads(2)
    _init(rank, world, path)
    batch = O.structured_batch(GLOBAL_B, 0, T, S)
    m = _step(batch, P.shard_bounds(GLOBAL_B, rank, world), P.GradSync())
    torch.save({"train": m._engine.train.data.clone(), "grad": m._engine.grad.clone()}, f"{out}.{rank}")
    dist.destroy_process_group()


def _worker_syncbn(rank, world, path, out):
    from cstp_b200 import parallel as P
    from oracle import cstp_oracle as O
    torch.set_num_threads(2)
    _init(rank, world, path)
    batch = O.structured_batch(GLOBAL_B, 0, T, S)
    m = _step(batch, P.shard_bounds(GLOBAL_B, rank, world), P.GradSync(), P.BnSync())
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    torch.save({"grad": m._engine.grad.clone(), "losses": m._engine.losses.clone(), "state": sd}, f"{out}.{rank}")
    dist.destroy_process_group()


def _worker_config5(rank, world, path, out):
    """BASELINE config 5 semantics at toy size: SyncBN over the world + NT-Xent on the all-gathered projector outputs."""
    from cstp_b200 import parallel as P
    from oracle import cstp_oracle as O
    torch.set_num_threads(2)
    _init(rank, world, path)
    batch = O.structured_batch(GLOBAL_B, 0, T, S)
    m = _step(batch, P.shard_bounds(GLOBAL_B, rank, world), P.GradSync(), P.BnSync(),
              dict(weight=0.7, temperature=0.5, gather=True))
    torch.save({"grad": m._engine.grad.clone(), "losses": m._engine.losses.clone(), "ntxent": m._engine.ntxent_loss.clone(),
                "train": m._engine.train.data.clone()}, f"{out}.{rank}")
    dist.destroy_process_group()


def _spawn(fn, world=2):
    d = tempfile.mkdtemp()
    path, out = os.path.join(d, "rdzv"), os.path.join(d, "out")
    mp.spawn(fn, args=(world, path, out), nprocs=world, join=True)
    return [torch.load(f"{out}.{r}", weights_only=False) for r in range(world)]


def test_collectives_world2():
    assert all(_spawn(_worker_collectives))


def test_data_parallel_step_world2():
    from cstp_b200 import engine
    from oracle import cstp_oracle as O
    r0, r1 = _spawn(_worker_step)
    assert torch.equal(r0["train"], r1["train"])            # replicas never drift: identical bits after the update
    assert torch.equal(r0["grad"], r1["grad"])
    # single-process reference: average of the two shards' gradients
    saved = engine.ops, engine.ACT_DTYPE
    try:
        batch = O.structured_batch(GLOBAL_B, 0, T, S)
        grads = []

        def capture(flat):
            grads.append(flat.clone())
        _step(batch, (0, 2), capture)
        _step(batch, (2, 4), capture)
    finally:
        engine.ops, engine.ACT_DTYPE = saved
    mean = (grads[0] + grads[1]) / 2
    # different thread counts -> different fp32 summation order (and the odd ReLU tie): norm-wise comparison
    assert ((r0["grad"] - mean).norm() / mean.norm()).item() < 5e-3


def test_world_synchronised_batchnorm_equals_global_batch():
    """bn_sync=BnSync(): two ranks x 2 samples with cross-rank BatchNorm statistics == one process x 4 samples
    (gradients after the data-parallel mean, BatchNorm running statistics, updated weights)."""
    from cstp_b200 import engine
    from oracle import cstp_oracle as O
    r0, r1 = _spawn(_worker_syncbn)
    assert torch.equal(r0["grad"], r1["grad"])
    saved = engine.ops, engine.ACT_DTYPE
    try:
        single = _step(O.structured_batch(GLOBAL_B, 0, T, S), (0, GLOBAL_B))
        g1 = single._engine.grad.clone()
        s1 = {k: v.clone() for k, v in single.state_dict().items()}
        l1 = single._engine.losses.clone()
    finally:
        engine.ops, engine.ACT_DTYPE = saved
    # global loss = mean of the per-rank losses (equal shard sizes)
    assert torch.allclose((r0["losses"] + r1["losses"]) / 2, l1, rtol=1e-4, atol=1e-5)
    assert ((r0["grad"] - g1).norm() / g1.norm()).item() < 1e-2          # fp32 ReLU ties / summation order only
    for k, v in s1.items():
        if "running" in k:
            assert torch.allclose(r0["state"][k], v, rtol=1e-4, atol=1e-6), k
    w = "online_net.conv3.block1.conv1.spatial_conv.weight"
    assert ((r0["state"][w] - s1[w]).norm() / s1[w].norm()).item() < 1e-4


def test_config5_syncbn_and_gathered_ntxent_equal_global_batch():
    """Two ranks x 2 samples with world BatchNorm statistics and NT-Xent over the all-gathered embeddings (global
    negatives) == one process x 4 samples: same NT-Xent loss on every rank, same gradients after the data-parallel mean."""
    from cstp_b200 import engine
    from oracle import cstp_oracle as O
    r0, r1 = _spawn(_worker_config5)
    assert torch.equal(r0["grad"], r1["grad"]) and torch.equal(r0["train"], r1["train"])
    assert torch.equal(r0["ntxent"], r1["ntxent"])                 # every rank evaluates the same global loss
    saved = engine.ops, engine.ACT_DTYPE
    try:
        single = _step(O.structured_batch(GLOBAL_B, 0, T, S), (0, GLOBAL_B), ntxent=dict(weight=0.7, temperature=0.5))
        g1, nx1, l1 = single._engine.grad.clone(), single._engine.ntxent_loss.clone(), single._engine.losses.clone()
        plain = _step(O.structured_batch(GLOBAL_B, 0, T, S), (0, GLOBAL_B))
        g0 = plain._engine.grad.clone()
    finally:
        engine.ops, engine.ACT_DTYPE = saved
    assert torch.allclose(r0["ntxent"], nx1, rtol=1e-4, atol=1e-5) and nx1.item() > 0
    assert torch.allclose((r0["losses"] + r1["losses"]) / 2, l1, rtol=1e-4, atol=1e-5)
    assert ((r0["grad"] - g1).norm() / g1.norm()).item() < 1e-2
    assert ((g1 - g0).norm() / g0.norm()).item() > 1e-2             # the NT-Xent term really reaches the backbone gradients


@pytest.mark.parametrize("gb,world", [(128, 8), (60, 6), (60, 8), (4, 2)])
def test_sharding_matches_reference_loader(gb, world):
    """utils.py:111: per-rank batch = int(batch_size / world_size); shards are contiguous and disjoint."""
    from cstp_b200 import parallel as P
    b = P.per_rank_batch(gb, world)
    assert b == int(gb / world)
    spans = [P.shard_bounds(gb, r, world) for r in range(world)]
    assert all(hi - lo == b for lo, hi in spans)
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1)) and spans[-1][1] <= gb
