"""Generates tests/golden/clips_ref.npz + clips_trace.json from the UNMODIFIED reference data pipeline.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden_clips.py

The reference classes (data_process/datasets.py UcfRepreBYOLSpPre / Kin400RepreLMDB, data_process/preprocess_data.py
get_transforms('pre_train')) are imported as they are; `decord` and `lmdb` (absent here, only used by other loaders) are
stubbed as empty modules, frames come from oracle.clip_oracle.synthetic_video through a patched `Image.open` / LMDB
record, and thin logging wrappers around Pillow / torchvision entry points record every decision the reference takes:
frames opened, 90-degree transposes, crop boxes, rotation angles, colour-jitter calls, gray channels, blur radii, flips.

  clips_trace.json : per case {seed, variant, total_frames, w, h, labels, trace of both clips}   (all cases)
  clips_ref.npz    : the two output clips of a few cases as uint8 ((x + 1) * 127.5 is exact for ToTensor + 'tf' output)
"""
from __future__ import annotations

import io
import json
import os
import random
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CSTP_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.clip_oracle import synthetic_video  # noqa: E402

for m in ("decord", "lmdb"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.modules["decord"].VideoReader = object
sys.modules["decord"].cpu = lambda *a, **k: None
sys.path.insert(0, REF)
from data_process import datasets as ds  # noqa: E402
from data_process import preprocess_data as pp  # noqa: E402
from PIL import Image, ImageFilter  # noqa: E402

TRACE: list = []


def _install_tracing():
    img_cls = Image.Image
    o_crop, o_rot, o_tr, o_filter = img_cls.crop, img_cls.rotate, img_cls.transpose, img_cls.filter

    def crop(self, box=None):
        TRACE.append(("crop", [int(v) for v in box]))
        return o_crop(self, box)

    def rotate(self, angle, *a, **k):
        TRACE.append(("rotate", float(angle)))
        return o_rot(self, angle, *a, **k)

    def transpose(self, method):
        TRACE.append(("transpose", int(method)))
        return o_tr(self, method)

    def filt(self, f):
        if isinstance(f, ImageFilter.GaussianBlur):
            TRACE.append(("blur", float(f.radius)))
        return o_filter(self, f)

    img_cls.crop, img_cls.rotate, img_cls.transpose, img_cls.filter = crop, rotate, transpose, filt
    for name in ("brightness", "contrast", "saturation", "hue"):
        orig = getattr(pp.F, "adjust_" + name)

        def wrapped(img, factor, _o=orig, _n=name):
            TRACE.append((_n, float(factor)))
            return _o(img, factor)
        setattr(pp.F, "adjust_" + name, wrapped)
    o_choice = np.random.choice

    def choice(*a, **k):
        r = o_choice(*a, **k)
        TRACE.append(("gray", int(r)))
        return r
    np.random.choice = choice


class _Opts:
    sample_duration = 16
    sample_size = 112
    task = "loss_com"


class _Txn:
    def __init__(self, raw):
        self.raw = raw

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def get(self, key):
        return self.raw


class _Env:
    def __init__(self, raw):
        self.raw = raw

    def begin(self, write=False):
        return _Txn(self.raw)


def run_case(seed: int, variant: str, total_frames: int, w: int, h: int):
    video = synthetic_video(total_frames + 1, w, h, seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    sp = pp.get_transforms("pre_train", _Opts)
    TRACE.clear()
    if variant == "ucf":
        obj = object.__new__(ds.UcfRepreBYOLSpPre)
        obj.opts, obj.sp_transform, obj.data_type = _Opts, sp, "train"
        obj.data = [("/video", "0", total_frames)]
        o_open = ds.Image.open

        def fake_open(path, *a, **k):
            n = int(os.path.basename(path).split(".")[0])
            TRACE.append(("open", n))
            return Image.fromarray(video[n - 1], "RGB")          # file 00001.jpg is video[0]
        ds.Image.open = fake_open
        try:
            clips, labels = obj[0]
        finally:
            ds.Image.open = o_open
    else:
        obj = object.__new__(ds.Kin400RepreLMDB)
        obj.opts, obj.sp_transform, obj.data_type = _Opts, sp, "train"
        obj.data = [("video", b"000000000", 0, total_frames)]

        class _Raw(list):
            def __getitem__(self, i):
                TRACE.append(("open", int(i)))
                return list.__getitem__(self, i)
        frames = _Raw([video[i] for i in range(total_frames)])
        obj.env = _Env(frames)
        o_loads = ds.msgpack.loads
        ds.msgpack.loads = lambda raw, *a, **k: raw
        obj.pil_from_raw_rgb = lambda arr: Image.fromarray(arr, "RGB")
        try:
            clips, labels = obj[0]
        finally:
            ds.msgpack.loads = o_loads
    spa, tem, pb, rot = labels
    return clips, dict(seed=seed, variant=variant, total_frames=total_frames, w=w, h=h,
                       labels=[int(spa), int(tem), int(pb), [int(rot[0]), int(rot[1])]], trace=[list(t) for t in TRACE])


def main():
    import contextlib
    _install_tracing()
    cases, pixels = [], {}
    shapes = [(320, 240), (171, 128), (160, 120), (128, 171)]
    totals = [10, 16, 29, 31, 64, 90, 150, 300]
    seed = 0
    with contextlib.redirect_stdout(io.StringIO()):           # get_transforms prints the transform on every call
        for variant in ("ucf", "kinetics"):
            for total in totals:
                for (w, h) in shapes:
                    for rep in range(4):
                        seed += 1
                        clips, info = run_case(seed, variant, total, w, h)
                        cases.append(info)
        # pixel fixtures: a few cases (null + base chains, long and short videos); smaller frames keep the file small
        want = [("ucf", 150, 160, 120), ("ucf", 64, 171, 128), ("ucf", 10, 160, 120), ("kinetics", 90, 160, 120)]
        pix_cases = []
        seed = 1000
        need_base = 3
        for variant, total, w, h in want:
            for rep in range(3):
                seed += 1
                clips, info = run_case(seed, variant, total, w, h)
                has_base = any(t[0] == "rotate" for t in info["trace"])
                if rep == 0 or (has_base and need_base > 0):
                    need_base -= 1 if has_base else 0
                    key = "case%d" % len(pix_cases)
                    for v in range(2):
                        u8 = torch.round((clips[v] + 1.0) * 127.5).to(torch.uint8)
                        assert torch.equal(u8.float() / 255 * 2.0 - 1.0, clips[v]), "output is not an exact uint8 image"
                        pixels["%s_view%d" % (key, v)] = u8.numpy()
                    pix_cases.append(info)
    out = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(out, "clips_trace.json"), "w") as f:
        json.dump(dict(cases=cases, pixel_cases=pix_cases, pillow=Image.__version__), f, separators=(",", ":"))
    np.savez_compressed(os.path.join(out, "clips_ref.npz"), **pixels)
    print("cases:", len(cases), "pixel cases:", len(pix_cases),
          "base-chain pixel cases:", sum(any(t[0] == "rotate" for t in c["trace"]) for c in pix_cases))


if __name__ == "__main__":
    main()
