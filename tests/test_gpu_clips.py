"""SURVEY.md 8(f-2) on the GPU: csrc/clip_pipeline.cu through the C ABI (`cstp_clip_assemble`) against
  * the clips the UNMODIFIED reference pipeline produced for the golden cases (tests/golden/clips_ref.npz), and
  * the Pillow pixel oracle (oracle/clip_oracle.py) on freshly seeded plans, incl. forced base-transform chains, every
    rotation label, crops reaching outside the frame, and a batch mixing frame sizes.
The bar is bit-exact (the whole pipeline is integer / exactly-rounded arithmetic)."""
import os
import random

import numpy as np
import pytest
import torch

from cstp_b200.data_process.clip_plan import PretrainClipSampler
from cstp_b200.data_process.gpu_clips import GpuClipPipeline, collate_labels
from oracle.clip_oracle import render_plan, synthetic_video
from tests.test_clip_pipeline import GOLD, TRACES, _force_base, seeded_plan

pytestmark = pytest.mark.gpu


def _video(case):
    n = case["total_frames"]
    return synthetic_video(n + 1, case["w"], case["h"], case["seed"])


def test_golden_reference_clips_bit_exact():
    ref = np.load(os.path.join(GOLD, "clips_ref.npz"))
    cases = TRACES["pixel_cases"]
    plans = [seeded_plan(c) for c in cases]
    videos = [torch.from_numpy(_video(c)).cuda() for c in cases]
    x1, x2 = GpuClipPipeline().assemble(plans, videos)
    torch.cuda.synchronize()
    for i in range(len(cases)):
        for v, x in enumerate((x1, x2)):
            want = torch.from_numpy(ref["case%d_view%d" % (i, v)]).float() / 255 * 2.0 - 1.0
            got = x[i].cpu()
            assert torch.equal(got, want), (i, v, (got - want).abs().max().item(), (got != want).sum().item())


@pytest.mark.parametrize("seed", [11, 12])
def test_seeded_batches_match_pixel_oracle(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    shapes = [(320, 240), (171, 128), (128, 171), (160, 120), (340, 256), (112, 112)]
    plans, vids = [], []
    for b in range(8):
        w, h = shapes[b % len(shapes)]
        total = [12, 40, 100][b % 3]
        variant = "ucf" if b % 2 == 0 else "kinetics"
        plan = PretrainClipSampler(16, 112, variant).plan(total, w, h)
        if b % 2 == 1:
            plan = _force_base(plan, seed * 100 + b)
        if b == 3:                                   # a crop that reaches outside the rotated frame (Pillow pads black)
            v = plan.views[1]
            v.box = (v.box[0] - 9, v.box[1] - 5, v.box[2] - 9, v.box[3] - 5)
        plans.append(plan)
        vids.append(synthetic_video(total + 1, w, h, seed + b))
    pipe = GpuClipPipeline()
    x1, x2 = pipe.assemble(plans, [torch.from_numpy(v).cuda() for v in vids])
    y1, y2 = pipe.assemble(plans, [torch.from_numpy(v).cuda() for v in vids])      # staging reuse: same answer
    torch.cuda.synchronize()
    assert torch.equal(x1, y1) and torch.equal(x2, y2)
    for b, (plan, vid) in enumerate(zip(plans, vids)):
        want = render_plan(plan, vid)
        for v, x in enumerate((x1, x2)):
            got = x[b].cpu()
            assert torch.equal(got, want[v]), (seed, b, v, (got - want[v]).abs().max().item(), (got != want[v]).sum().item())
    spa, tem, pb, (r1, r2) = collate_labels(plans, device="cuda")
    assert spa.tolist() == [p.spa_label for p in plans] and r2.tolist() == [p.rot_labels[1] for p in plans]


def test_assembled_clips_drive_a_training_step():
    """The clips and labels go straight into the fused pretraining step (main_byol.py:43-49 hand-off)."""
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    random.seed(5)
    np.random.seed(5)
    torch.manual_seed(5)
    plans, vids = [], []
    for b in range(2):
        plans.append(PretrainClipSampler().plan(60, 160, 120))
        vids.append(torch.from_numpy(synthetic_video(61, 160, 120, b)).cuda())
    x1, x2 = GpuClipPipeline().assemble(plans, vids)
    spa, tem, pb, (r1, r2) = collate_labels(plans, device="cuda")
    torch.manual_seed(1)
    model = R21DBYOL(pretrain=True).cuda()
    losses = model.train_step(x1, x2, (spa, tem, pb, r1, r2), (0.1, 1, 1, 1, 1), lr=0.03)
    torch.cuda.synchronize()
    assert torch.isfinite(losses).all()


def test_batch_source_feeds_the_epoch_driver():
    from types import SimpleNamespace
    from cstp_b200.data_process.datasets import GpuVideoStore, PretrainBatches
    from cstp_b200.models.pace.r21d_byol import R21DBYOL
    from cstp_b200.train import pretrain_epochs
    random.seed(3)
    np.random.seed(3)
    torch.manual_seed(3)
    store = GpuVideoStore()
    totals = [40, 90, 20, 64, 33]
    for i, n in enumerate(totals):
        store.put(i, synthetic_video(n, 160, 120, i))
    batches = PretrainBatches(store, totals, batch_size=2, variant="ucf")
    got = list(batches(1))
    assert len(got) == 2 and got[0][0].shape == (2, 3, 16, 112, 112) and got[0][2][0].dtype == torch.int64
    ref = PretrainBatches(store, totals, batch_size=2, reference_format=True)
    clips, targets = next(iter(ref(1)))
    assert len(clips) == 2 and len(targets) == 4 and len(targets[3]) == 2
    torch.manual_seed(1)
    model = R21DBYOL(pretrain=True).cuda()
    opts = SimpleNamespace(n_epochs=1, learning_rate=0.03, momentum=0.9, weight_decay=5e-4, loss_weight=[0.1, 1, 1, 1, 1],
                           clip_grad_norm=1)
    rows = pretrain_epochs(model, batches, opts)
    assert len(rows) == 1 and np.isfinite(rows[0]["loss"])


def test_finetune_val_test_clips_bit_exact():
    """`assemble_clips` on UcfFineTune-style plans (random-sized crop + colour jitter, scale + centre crop, the multi-clip
    test windows) against the reference's own clips and the Pillow oracle."""
    from oracle.clip_oracle import render_view
    from tests.test_clip_pipeline import FT_TRACES, seeded_ft_plans
    ref = np.load(os.path.join(GOLD, "clips_ft_ref.npz"))
    pipe = GpuClipPipeline()
    for i, case in enumerate(FT_TRACES["pixel_cases"]):
        plans = seeded_ft_plans(case)
        video = synthetic_video(case["total_frames"] + 1, case["w"], case["h"], case["seed"])
        dvid = torch.from_numpy(video).cuda()
        out = pipe.assemble_clips(plans, [dvid] * len(plans))
        torch.cuda.synchronize()
        assert out.shape == (len(plans), 3, 16, 112, 112)
        for j in case["kept"]:
            want = torch.from_numpy(ref["case%d_clip%d" % (i, j)]).float() / 255 * 2.0 - 1.0
            assert torch.equal(out[j].cpu(), want), (i, j)
        for j, p in enumerate(plans):                          # every clip of the case against the oracle
            assert torch.equal(out[j].cpu(), render_view(p.view, video, p.frame_base)), (i, j)
