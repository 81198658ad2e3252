"""Generates tests/golden/clips_ft_trace.json + clips_ft_ref.npz from the UNMODIFIED reference finetune data pipeline
(data_process/datasets.py UcfFineTune with get_transforms('img' | 'img_val' | 'img_test')), the same way
oracle/make_golden_clips.py does for the pretraining pipeline (run in the build container only):

    python oracle/make_golden_clips_ft.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden_clips as G  # noqa: E402  (imports the reference with decord / lmdb stubbed)
from oracle.clip_oracle import synthetic_video  # noqa: E402
from PIL import Image  # noqa: E402

ds, pp, TRACE = G.ds, G.pp, G.TRACE


class _Opts:
    sample_duration = 16
    sample_size = 112
    task = "ft_all"
    pb_rate = 4


def _trace_resize():
    o_resize = Image.Image.resize

    def resize(self, size, *a, **k):
        TRACE.append(("resize", [int(size[0]), int(size[1])]))
        return o_resize(self, size, *a, **k)
    Image.Image.resize = resize


def run_case(seed: int, mode: str, total_frames: int, w: int, h: int):
    video = synthetic_video(total_frames + 1, w, h, seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    sp = pp.get_transforms({"train": "img", "val": "img_val", "test": "img_test"}[mode], _Opts)
    TRACE.clear()
    obj = object.__new__(ds.UcfFineTune)
    obj.opts, obj.sp_transform, obj.data_type = _Opts, sp, mode
    obj.data = [("/video", 7, total_frames)]
    o_open = ds.Image.open

    def fake_open(path, *a, **k):
        n = int(os.path.basename(path).split(".")[0])
        TRACE.append(("open", n))
        return Image.fromarray(video[n - 1], "RGB")
    ds.Image.open = fake_open
    try:
        clips, label = obj[0]
    finally:
        ds.Image.open = o_open
    clips = torch.as_tensor(clips)
    if mode != "test":
        clips = clips[None]                                   # (n_clips, 3, T, S, S)
    return clips, dict(seed=seed, mode=mode, total_frames=total_frames, w=w, h=h, label=int(label),
                       n_clips=int(clips.shape[0]), trace=[list(t) for t in TRACE])


def main():
    G._install_tracing()
    _trace_resize()
    cases, pixels, pix_cases = [], {}, []
    seed = 5000
    with contextlib.redirect_stdout(io.StringIO()):
        for mode in ("train", "val", "test"):
            for total in (10, 40, 61, 150):
                for (w, h) in ((320, 240), (171, 128), (128, 171), (160, 120)):
                    for rep in range(3 if mode == "train" else 1):
                        seed += 1
                        _, info = run_case(seed, mode, total, w, h)
                        cases.append(info)
        want = [("train", 64, 160, 120), ("train", 64, 171, 128), ("val", 40, 171, 128), ("test", 100, 160, 120)]
        seed = 6000
        for mode, total, w, h in want:
            for rep in range(6):
                seed += 1
                clips, info = run_case(seed, mode, total, w, h)
                jit = any(t[0] in ("brightness", "hue") for t in info["trace"])
                if mode != "train" or (jit if (mode, w) == ("train", 160) else not jit):
                    key = "case%d" % len(pix_cases)
                    keep = [0] if clips.shape[0] == 1 else [0, clips.shape[0] - 1]
                    for j in keep:
                        u8 = torch.round((clips[j] + 1.0) * 127.5).to(torch.uint8)
                        assert torch.equal(u8.float() / 255 * 2.0 - 1.0, clips[j].float())
                        pixels["%s_clip%d" % (key, j)] = u8.numpy()
                    info["kept"] = keep
                    pix_cases.append(info)
                    break
    out = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(out, "clips_ft_trace.json"), "w") as f:
        json.dump(dict(cases=cases, pixel_cases=pix_cases, pillow=Image.__version__), f, separators=(",", ":"))
    np.savez_compressed(os.path.join(out, "clips_ft_ref.npz"), **pixels)
    print("cases:", len(cases), "pixel cases:", [(c["mode"], c["n_clips"], c["kept"]) for c in pix_cases])


if __name__ == "__main__":
    main()
