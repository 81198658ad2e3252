"""the pretraining clip pipeline.  CPU part:
  * cstp_b200.data_process.clip_plan.PretrainClipSampler against traces of the UNMODIFIED reference classes
    (tests/golden/clips_trace.json, produced by oracle/make_golden_clips.py): labels, frame numbers, rotation codes, crop
    boxes, flips and every base-transform parameter bit-exact, for 256 seeded (variant, video length, frame size) cases;
  * oracle.clip_oracle.render_plan (the pixel oracle the GPU tests use) against the reference's output clips: bit-exact.
"""
import json
import os
import random

import numpy as np
import pytest
import torch

from cstp_b200.data_process.clip_plan import PretrainClipSampler
from oracle.clip_oracle import ROT_METHOD, render_plan, synthetic_video

GOLD = os.path.join(os.path.dirname(__file__), "golden")
with open(os.path.join(GOLD, "clips_trace.json")) as f:
    TRACES = json.load(f)


def seeded_plan(case):
    random.seed(case["seed"])
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    return PretrainClipSampler(16, 112, case["variant"]).plan(case["total_frames"], case["w"], case["h"])


def dedupe(events):
    out = []
    for e in events:
        e = tuple(e) if not isinstance(e[1], list) else (e[0], tuple(e[1]))
        if out and out[-1] == e and e[0] not in ("open", "gray"):
            continue
        out.append(e)
    return out


def expected_events(plan):
    """The Pillow / torchvision call sequence the reference makes for this plan (one entry per frame, as traced)."""
    v1, v2 = plan.views
    ev = []
    m1, m2 = ROT_METHOD[v1.rot], ROT_METHOD[v2.rot]
    short = plan.draws.get("short", None)
    if short:
        for n in v1.frames:                      # datasets.py:893-910: one open, both rotations
            ev.append(("open", n))
            if m1 is not None:
                ev.append(("transpose", int(m1)))
            if m2 is not None:
                ev.append(("transpose", int(m2)))
    else:
        for v, m in ((v1, m1), (v2, m2)):
            for n in v.frames:
                ev.append(("open", n))
                if m is not None:
                    ev.append(("transpose", int(m)))
    for v in (v1, v2):
        T = len(v.frames)
        ev += [("crop", tuple(v.box))] * T
        if v.base:
            ev += [("rotate", v.angle)] * T
            if v.jitter is not None:
                # transforms.Compose of the shuffled ops is applied frame by frame: op order cycles per frame
                for _ in range(T):
                    ev += [(name, factor) for name, factor in v.jitter]
            if v.gray is not None:
                ev += [("gray", c) for c in v.gray]
            if v.blur_sigma is not None:
                ev += [("blur", v.blur_sigma)] * T
        if v.flip:
            ev += [("transpose", 0)] * T
    return ev


@pytest.mark.parametrize("which", ["cases", "pixel_cases"])
def test_sampler_matches_reference_traces(which):
    n_base = n_short = n_retry = 0
    for case in TRACES[which]:
        plan = seeded_plan(case)
        assert plan.labels() == case["labels"], case
        got = dedupe(expected_events(plan))
        want = dedupe(case["trace"])
        assert got == want, (case["seed"], case["variant"], case["total_frames"], got[:40], want[:40])
        n_base += plan.views[0].base + plan.views[1].base
        n_short += bool(plan.draws.get("short"))
        n_retry += plan.draws["temporal_retries"] > 0
    if which == "cases":          # the fixture exercises every branch
        assert n_base > 50 and n_short > 10 and n_retry > 20


def test_label_and_index_ranges():
    random.seed(7)
    np.random.seed(7)
    torch.manual_seed(7)
    for variant in ("ucf", "kinetics"):
        s = PretrainClipSampler(16, 112, variant)
        for total in (15, 40, 200):
            for _ in range(50):
                p = s.plan(total, 320, 240)
                assert 0 <= p.spa_label <= 4 and 0 <= p.tem_label <= 4 and 0 <= p.pb_label <= 3
                assert all(0 <= r <= 3 for r in p.rot_labels)
                lo, hi = (1, total) if variant == "ucf" else (0, total - 1)
                for v in p.views:
                    assert len(v.frames) == 16 and all(lo <= n <= hi for n in v.frames), (variant, total, v.frames)
                    x0, y0, x1, y1 = v.box
                    assert x1 - x0 == p.views[0].box[2] - p.views[0].box[0]      # the second crop keeps the first one's size
                    assert y1 - y0 == p.views[0].box[3] - p.views[0].box[1]


def test_pixel_oracle_matches_reference_clips():
    ref = np.load(os.path.join(GOLD, "clips_ref.npz"))
    for i, case in enumerate(TRACES["pixel_cases"]):
        plan = seeded_plan(case)
        n = case["total_frames"]
        video = synthetic_video(n + 1, case["w"], case["h"], case["seed"])
        if case["variant"] == "kinetics":
            video = video[:n]
        clips = render_plan(plan, video)
        for v in range(2):
            want = torch.from_numpy(ref["case%d_view%d" % (i, v)]).float() / 255 * 2.0 - 1.0
            assert clips[v].shape == (3, 16, 112, 112)
            assert torch.equal(clips[v], want), (i, v, (clips[v] - want).abs().max().item())


# ---------------------------------------------------------------------------------------------------------------
# host side of the GPU path: descriptors + integer tables, executed by the numpy stand-in of csrc/clip_pipeline.cu
def _emulate(plan, video, S=112):
    from cstp_b200.data_process.gpu_clips import compile_view
    from tests.emulate_clips import run_view
    H, W = video.shape[1:3]
    return [torch.from_numpy(run_view(*compile_view(v, plan.frame_base, W, H, S)[:2], video, len(v.frames), S))
            for v in plan.views]


def _force_base(plan, seed):
    """Give both views a full base chain (the sampler only takes it 30 % of the time)."""
    rng = random.Random(seed)
    for v in plan.views:
        v.base = True
        v.angle = rng.uniform(-10, 10)
        ops = [("brightness", rng.uniform(0.6, 1.4)), ("contrast", rng.uniform(0.6, 1.4)),
               ("saturation", rng.uniform(0.6, 1.4)), ("hue", rng.uniform(-0.1, 0.1))]
        rng.shuffle(ops)
        v.jitter = ops
        v.gray = [rng.randrange(3) for _ in v.frames] if rng.random() < 0.5 else None
        v.blur_sigma = rng.uniform(0.1, 2.0)
    return plan


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_descriptor_path_matches_pixel_oracle(idx):
    """plan -> cstp_clip_view descriptors -> integer arithmetic == Pillow, bit-exact (null and base chains, all four
    rotation labels, crops reaching outside the frame)."""
    case = TRACES["pixel_cases"][idx]
    plan = seeded_plan(case)
    if idx % 2 == 1:
        plan = _force_base(plan, idx)
    plan.views[0].rot, plan.views[1].rot = idx % 4, (idx + 1) % 4       # pixel path only: any rotation pair is legal
    w, h = case["w"], case["h"]
    for v in plan.views:                                               # keep the boxes inside / around the rotated frame
        wr, hr = (h, w) if v.rot & 1 else (w, h)
        bw, bh = min(v.box[2] - v.box[0], wr), min(v.box[3] - v.box[1], hr)
        v.box = (wr - bw + 3 * (idx % 2), 0 - 2 * (idx % 2), wr + 3 * (idx % 2), bh - 2 * (idx % 2))
    n = case["total_frames"]
    video = synthetic_video(n + 1, w, h, case["seed"])
    # two frames per clip are enough for the CPU stand-in (every frame runs the same arithmetic)
    for v in plan.views:
        v.frames = v.frames[:2]
        v.gray = v.gray[:2] if v.gray is not None else None
    want = render_plan(plan, video)
    got = _emulate(plan, video)
    for v in range(2):
        assert torch.equal(got[v], want[v]), (idx, v, (got[v] - want[v]).abs().max().item())


def test_resample_tables_cover_every_crop_size():
    from cstp_b200.data_process.gpu_clips import KMAX, resample_tables
    for n in (1, 2, 37, 111, 112, 113, 240, 320, 616):
        t = resample_tables(n, 112)
        assert t.shape == (112, 2 + KMAX) and (t[:, 1] > 0).all() and (t[:, 0] + t[:, 1] <= n).all()
        s = t[:, 2:].astype(np.int64).sum(1)
        assert np.abs(s - (1 << 22)).max() <= KMAX            # taps sum to 1.0 in Q22 up to rounding
    with pytest.raises(ValueError):
        resample_tables(700, 112)


def test_collate_labels_layout():
    from cstp_b200.data_process.gpu_clips import collate_labels
    plans = [seeded_plan(c) for c in TRACES["cases"][:5]]
    spa, tem, pb, (r1, r2) = collate_labels(plans)
    for i, c in enumerate(TRACES["cases"][:5]):
        assert [int(spa[i]), int(tem[i]), int(pb[i]), [int(r1[i]), int(r2[i])]] == c["labels"]
    assert spa.dtype == torch.int64 and spa.shape == (5,)


def test_pipeline_has_no_cpu_fallback():
    from cstp_b200.data_process.gpu_clips import GpuClipPipeline
    from cstp_b200.lib import CstpError
    with pytest.raises(CstpError):
        GpuClipPipeline(device="cpu").assemble([], [])


def test_index_order_is_the_distributed_samplers():
    from torch.utils.data import DistributedSampler
    from cstp_b200.data_process.datasets import distributed_indices
    data = list(range(103))
    for world in (1, 2, 4):
        for rank in range(world):
            s = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=0)
            for epoch in (0, 3):
                s.set_epoch(epoch)
                assert list(iter(s)) == distributed_indices(len(data), epoch, rank, world, 0, True)


# ---------------------------------------------------------------------------------------------------------------
# finetune / validation / test clips (UcfFineTune + get_transforms('img' | 'img_val' | 'img_test'))
with open(os.path.join(GOLD, "clips_ft_trace.json")) as f:
    FT_TRACES = json.load(f)


def seeded_ft_plans(case):
    from cstp_b200.data_process.clip_plan import FinetuneClipSampler
    random.seed(case["seed"])
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    s = FinetuneClipSampler(16, 112, 4)
    if case["mode"] == "train":
        return [s.plan_train(case["total_frames"], case["w"], case["h"])]
    if case["mode"] == "val":
        return [s.plan_val(case["total_frames"], case["w"], case["h"])]
    return s.plan_test(case["total_frames"], case["w"], case["h"])


def expected_ft_events(plans, w, h):
    ev = []
    for p in plans:
        v, T = p.view, len(p.view.frames)
        ev += [("open", n) for n in v.frames]
        if v.resize is None:
            ev += [("crop", tuple(v.box))] * T + [("resize", (112, 112))] * T
        else:
            if tuple(v.resize) != (w, h):
                ev += [("resize", tuple(v.resize))] * T
            ev += [("crop", tuple(v.box))] * T
        if v.jitter is not None:
            for _ in range(T):
                ev += [(name, factor) for name, factor in v.jitter]
    return ev


@pytest.mark.parametrize("which", ["cases", "pixel_cases"])
def test_finetune_sampler_matches_reference_traces(which):
    n_jit = n_multi = n_fallback = 0
    for case in FT_TRACES[which]:
        plans = seeded_ft_plans(case)
        assert len(plans) == case["n_clips"], case
        assert dedupe(expected_ft_events(plans, case["w"], case["h"])) == dedupe(case["trace"]), \
            (case["seed"], case["mode"], case["total_frames"], case["w"], case["h"])
        n_jit += plans[0].view.jitter is not None
        n_multi += len(plans) > 1
        n_fallback += case["mode"] == "train" and plans[0].view.resize is not None
    if which == "cases":
        assert n_jit >= 5 and n_multi >= 8


def test_finetune_pixel_oracle_and_descriptor_path_match_reference_clips():
    """Pillow oracle == the reference's finetune / val / test clips, and the integer descriptor path == both (resize views:
    the whole frame through slices of the tap tables)."""
    from oracle.clip_oracle import render_view
    ref = np.load(os.path.join(GOLD, "clips_ft_ref.npz"))
    for i, case in enumerate(FT_TRACES["pixel_cases"]):
        plans = seeded_ft_plans(case)
        video = synthetic_video(case["total_frames"] + 1, case["w"], case["h"], case["seed"])
        for j in case["kept"]:
            want = torch.from_numpy(ref["case%d_clip%d" % (i, j)]).float() / 255 * 2.0 - 1.0
            got = render_view(plans[j].view, video, plans[j].frame_base)
            assert torch.equal(got, want), (i, j, (got - want).abs().max().item())
            from cstp_b200.data_process.gpu_clips import compile_view
            from tests.emulate_clips import run_view
            v = plans[j].view
            keep = v.frames
            v.frames = v.frames[:2]                              # two frames are enough for the numpy stand-in
            d, coef, _ = compile_view(v, plans[j].frame_base, case["w"], case["h"], 112)
            emu = torch.from_numpy(run_view(d, coef, video, 2, 112))
            v.frames = keep
            assert torch.equal(emu, want[:, :2]), (i, j, (emu - want[:, :2]).abs().max().item())


def test_video_store_bookkeeping_on_cpu():
    """GpuVideoStore's residency logic does not need a GPU: capacity-bounded, oldest entries dropped first."""
    from cstp_b200.data_process.datasets import GpuVideoStore
    store = GpuVideoStore(device="cpu", capacity_bytes=3 * 4 * 6 * 5 * 3 + 10)
    for i in range(4):
        store.put(i, synthetic_video(4, 5, 6, i))
    assert len(store) == 3 and 0 not in store and 3 in store and store[3].shape == (4, 6, 5, 3)
    store.put(1, synthetic_video(4, 5, 6, 9))              # replacing an entry keeps the byte count right
    assert store.bytes == 3 * 4 * 6 * 5 * 3
    with pytest.raises(ValueError):
        store.put(7, np.zeros((2, 3, 4), np.uint8))
