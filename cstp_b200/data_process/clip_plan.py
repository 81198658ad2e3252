"""Host-side sampler of the pretraining clip pipeline: every random decision of one dataset item, taken in the
reference's draw order, returned as a plan the GPU kernels (cstp_b200/data_process/gpu_clips.py) execute.

Replaces the decision-taking half of (reference file:line)
  * data_process/datasets.py:859-948    UcfRepreBYOLSpPre.repre_train_clip   (variant="ucf": frame files are 1-based)
  * data_process/datasets.py:1308-1405  Kin400RepreLMDB.repre_train_clip     (variant="kinetics": 0-based frames; the
    second clip re-reads the first clip's frames, datasets.py:1397 -- reproduced, `clip2_reads_clip1=True`)
  * data_process/preprocess_data.py:771-780  TransformController (random.choices, weights [1, 0])
  * data_process/preprocess_data.py:713-741  TwoClipTransform (p = 0.3 for the base transform, per clip)
  * data_process/preprocess_data.py:479-565  ClipRandomSizedCropOverlap (first crop; second crop tied to the first by
    the spatial-overlap label)
  * data_process/preprocess_data.py:1112-1122 the base transform chain (RandomRotation(10), RandomApply(ColorJitter, 0.8),
    ClipRandomGray(0.2), RandomApply(GaussianBlur, 0.5), flip) and the null chain (flip)

The labels the pretext heads are trained on -- spa (overlap_spa), tem (overlap_tem), pb (playback rate), rot1 / rot2 --
and every index (frame numbers, crop boxes, rotation codes, flip flags) are BIT-EXACT with the reference when the three
generators it consumes (Python `random`, `numpy.random`, `torch` default generator) are seeded identically: the sampler
draws from the same generators with the same calls in the same order, and repeats the reference's double arithmetic
(e.g. `int((1 - 0.8) * 15) == 2`).  tests/test_clip_pipeline.py checks this against traces of the unmodified reference.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch

PACE = [1, 2, 4, 8]                                  # datasets.py:17
OVERLAP_TEM_RATE = [1.0, 0.8, 0.6, 0.4, 0.2]         # datasets.py:18
OVERLAP_SPA_RATE = [1.0, 0.8, 0.6, 0.4, 0.2]         # preprocess_data.py:18


@dataclass
class ViewPlan:
    frames: List[int]                  # the frame numbers the reference opens, in clip order
    rot: int                           # rotation label: the frame is turned by rot x 90 degrees counter-clockwise
    box: Tuple[int, int, int, int]     # crop (x0, y0, x1, y1) in the ROTATED frame; may reach outside (PIL pads black)
    base: bool                         # True: base transform chain, False: null chain (flip only)
    flip: bool
    angle: Optional[float] = None                          # RandomRotation angle (degrees, counter-clockwise)
    jitter: Optional[List[Tuple[str, float]]] = None       # colour-jitter ops in the order they are applied
    gray: Optional[List[int]] = None                       # per frame: the channel replicated into R, G and B
    blur_sigma: Optional[float] = None
    # finetune / test chains: the whole frame is first resized to (ow, oh) -- ClipScale -- and `box` then crops the RESIZED
    # frame (ClipCenterCrop); None: `box` crops the (rotated) source frame and the crop is resized to size x size
    resize: Optional[Tuple[int, int]] = None


@dataclass
class SamplePlan:
    views: Tuple[ViewPlan, ViewPlan]
    spa_label: int
    tem_label: int
    pb_label: int
    rot_labels: Tuple[int, int]
    frame_base: int                    # 1: frames are file numbers 00001.jpg.. (UCF); 0: list indices (Kinetics LMDB)
    draws: dict = field(default_factory=dict)      # bookkeeping (crop attempts, temporal retries)

    def labels(self):
        """[spa, tem, pb, [rot1, rot2]] -- what `__getitem__` returns next to the clips (datasets.py:856-857)."""
        return [self.spa_label, self.tem_label, self.pb_label, list(self.rot_labels)]


class PretrainClipSampler:
    """plan(total_frames, frame_w, frame_h) -> SamplePlan.  Consumes random / numpy.random / torch RNG state exactly as the
    reference's `__getitem__` does for a video of that length and frame size."""

    def __init__(self, sample_duration: int = 16, sample_size: int = 112, variant: str = "ucf", p_base: float = 0.3,
                 bottom_area: float = 0.2):
        if variant not in ("ucf", "kinetics"):
            raise ValueError("variant must be 'ucf' or 'kinetics'")
        self.T = int(sample_duration)
        self.size = int(sample_size)
        self.variant = variant
        self.p_base = p_base
        self.bottom_area = bottom_area

    # ------------------------------------------------------------------ temporal part (repre_train_clip)
    def _temporal(self, total_frames: int):
        T = self.T
        max_pb = int(np.log2(total_frames / (T - 1)))                              # datasets.py:872
        pb_label = random.randint(0, min(3, max_pb))
        rate = PACE[pb_label]
        clip_range = (T - 1) * rate
        rot1 = random.randint(0, 3)
        rot2 = random.randint(0, 3)
        base = 1 if self.variant == "ucf" else 0
        retries = 0
        if total_frames - clip_range <= 0:                                          # short video: wrap around, one clip twice
            idx, f = [], 0
            while len(idx) < T:
                idx.append(f)
                f += rate
                if f >= total_frames:
                    f = 0
            frames1 = [base + i for i in idx]
            frames2 = list(frames1)
            tem_label = 0
        else:
            if self.variant == "ucf":
                start = random.randint(1, total_frames - clip_range)                # datasets.py:913
            else:
                start = random.randint(0, total_frames - clip_range - 1)            # datasets.py:1364
            while True:
                tem_label = random.randint(0, 4)
                tem_rate = OVERLAP_TEM_RATE[tem_label]
                front_behind = random.randint(0, 1)
                shift = int((1 - tem_rate) * clip_range)                            # double arithmetic, truncation
                if front_behind == 0:
                    start2 = start - shift
                    if start2 < 1:
                        retries += 1
                        continue
                else:
                    start2 = start + shift
                    if start2 > total_frames - clip_range:
                        retries += 1
                        continue
                break
            offs = list(range(0, clip_range + 1, rate))
            frames1 = [start + i for i in offs]
            # datasets.py:1397 reads raw[start_frame + i] for the second clip as well
            frames2 = [(start if self.variant == "kinetics" else start2) + i for i in offs]
        return frames1, frames2, tem_label, pb_label, rot1, rot2, base, retries, total_frames - clip_range <= 0

    # ------------------------------------------------------------------ spatial part (ClipRandomSizedCropOverlap)
    def _crop_first(self, img_w: int, img_h: int):
        random.random()                          # `random.random() < self.threshold` with p = 1.0: drawn, always true
        attempts = 0
        while True:
            attempts += 1
            area = img_w * img_h
            target_area = random.uniform(self.bottom_area, 1) * area
            aspect = random.uniform(3. / 4, 4. / 3)
            w = int(round(math.sqrt(target_area * aspect)))
            h = int(round(math.sqrt(target_area / aspect)))
            if random.random() < 0.5:
                w, h = h, w
            if w <= img_w and h <= img_h:
                x1 = random.randint(0, img_w - w)
                y1 = random.randint(0, img_h - h)
                return (x1, y1, x1 + w, y1 + h), (w, h), (x1, y1), attempts

    def _crop_second(self, img_w: int, img_h: int, pick_size, pick_loc):
        random.random()
        p_w, p_h = pick_size
        p_x, p_y = pick_loc
        attempts = 0
        while True:
            attempts += 1
            random.uniform(self.bottom_area, 1)                 # target area and aspect ratio are drawn and not used
            random.uniform(3. / 4, 4. / 3)
            spa_label = random.randint(0, 4)
            spa_rate = OVERLAP_SPA_RATE[spa_label]
            corner = random.randint(0, 3)
            s_w = random.randint(int(spa_rate * p_w), p_w)
            s_h = int(spa_rate * p_w * p_h / s_w)
            if corner == 0:
                e_w, e_h = p_x + s_w, p_y + s_h
                ok = e_w - p_w >= 0 and e_h - p_h >= 0
            elif corner == 1:
                e_w, e_h = p_x + p_w - s_w + p_w, p_y + s_h
                ok = e_w <= img_w and e_h - p_h >= 0
            elif corner == 2:
                e_w, e_h = p_x + s_w, p_y + p_h - s_h + p_h
                ok = e_w - p_w >= 0 and e_h <= img_h
            else:
                e_w, e_h = p_x + p_w - s_w + p_w, p_y + p_h - s_h + p_h
                ok = e_w <= img_w and e_h <= img_h
            if ok:
                return (e_w - p_w, e_h - p_h, e_w, e_h), spa_label, attempts

    # ------------------------------------------------------------------ per-clip transform chains
    def _chain(self, view: ViewPlan):
        if view.base:
            view.angle = random.uniform(-10, 10)                                    # RandomRotation(10)
            if not (0.8 < torch.rand(1)):                                           # transforms.RandomApply(p=0.8)
                random.random()                                                     # ClipColorJitter p = 1.0: drawn
                # ClipColorJitter(0.4, 0.4, 0.4, 0.1): ranges [center - v, center + v] as _check_input builds them
                ops = [("brightness", random.uniform(1 - 0.4, 1 + 0.4)), ("contrast", random.uniform(1 - 0.4, 1 + 0.4)),
                       ("saturation", random.uniform(1 - 0.4, 1 + 0.4)), ("hue", random.uniform(0 - 0.1, 0 + 0.1))]
                random.shuffle(ops)
                view.jitter = ops
            if random.random() < 0.2:                                               # ClipRandomGray(p=0.2)
                view.gray = [int(np.random.choice(3)) for _ in range(self.T)]
            if not (0.5 < torch.rand(1)):                                           # RandomApply(GaussianBlur, p=0.5)
                view.blur_sigma = random.uniform(.1, 2.)                            # one draw per clip (idx % T == 0)
        view.flip = random.random() < 0.5

    # ------------------------------------------------------------------ one dataset item
    def plan(self, total_frames: int, frame_w: int, frame_h: int) -> SamplePlan:
        frames1, frames2, tem, pb, rot1, rot2, base, retries, short = self._temporal(total_frames)

        def rotated(rot):
            return (frame_h, frame_w) if rot in (1, 3) else (frame_w, frame_h)
        random.choices(range(2), weights=[1, 0])                                    # TransformController
        base1 = random.random() < self.p_base                                       # TwoClipTransform
        base2 = random.random() < self.p_base
        w1, h1 = rotated(rot1)
        box1, pick_size, pick_loc, a1 = self._crop_first(w1, h1)
        v1 = ViewPlan(frames1, rot1, box1, base1, False)
        self._chain(v1)
        w2, h2 = rotated(rot2)
        box2, spa, a2 = self._crop_second(w2, h2, pick_size, pick_loc)
        v2 = ViewPlan(frames2, rot2, box2, base2, False)
        self._chain(v2)
        return SamplePlan((v1, v2), spa, tem, pb, (rot1, rot2), base,
                          dict(temporal_retries=retries, crop1_attempts=a1, crop2_attempts=a2, short=short))


@dataclass
class ClipPlan:
    """One finetune / validation / test clip: a single view plus what the loader returns next to it."""
    view: ViewPlan
    frame_base: int = 1


class FinetuneClipSampler:
    """The clip decisions of `UcfFineTune` (data_process/datasets.py:951-1097) under get_transforms('img' | 'img_val' |
    'img_test') (data_process/preprocess_data.py:1131-1155), in the reference's draw order on Python `random`:

      train : start = randint(1, total - clip_range) (:1007-1017), then ClipRandomSizedCrop(size, bottom_area=0.2)
              (preprocess_data.py:440-476: random() < 1.0, up to ten (area, aspect, swap, x1, y1) attempts, else
              ClipScale(size) + ClipCenterCrop(size)) and ClipColorJitter(0.4, 0.4, 0.4, 0.1, p=0.3) (:659-663)
      val   : the same temporal draw (:1033-1047), ClipScale(short side 128 for size 112 / 256 for 224) + ClipCenterCrop
      test  : every window of (T - 1) * pb_rate frames plus the last one (:1062-1080), no random draw; same spatial chain
    """

    def __init__(self, sample_duration: int = 16, sample_size: int = 112, pb_rate: int = 4):
        self.T, self.size, self.rate = int(sample_duration), int(sample_size), int(pb_rate)
        self.short = {112: 128, 224: 256}.get(self.size)

    def _frames(self, total_frames: int) -> List[int]:
        clip_range = (self.T - 1) * self.rate
        if total_frames - clip_range <= 0:
            idx, f = [], 0
            while len(idx) < self.T:
                idx.append(f)
                f += self.rate
                if f >= total_frames:
                    f = 0
            return [1 + i for i in idx]
        start = random.randint(1, total_frames - clip_range)
        return [start + i for i in range(0, clip_range + 1, self.rate)]

    @staticmethod
    def _scale_center(w: int, h: int, short: int, size: int):
        """ClipScale(short) + ClipCenterCrop(size) (preprocess_data.py:843-866, 815-840) -> (resize or None, box)."""
        if (w <= h and w == short) or (h <= w and h == short):
            ow, oh = w, h
        elif w < h:
            ow, oh = short, int(short * h / w)
        else:
            ow, oh = int(short * w / h), short
        x1 = int(round((ow - size) / 2.))
        y1 = int(round((oh - size) / 2.))
        return (ow, oh), (x1, y1, x1 + size, y1 + size)

    def plan_train(self, total_frames: int, frame_w: int, frame_h: int) -> ClipPlan:
        frames = self._frames(total_frames)
        view = None
        random.random()                                        # ClipRandomSizedCrop p = 1.0: drawn, always true
        for _ in range(10):
            area = frame_w * frame_h
            target_area = random.uniform(0.2, 1) * area
            aspect = random.uniform(3. / 4, 4. / 3)
            w = int(round(math.sqrt(target_area * aspect)))
            h = int(round(math.sqrt(target_area / aspect)))
            if random.random() < 0.5:
                w, h = h, w
            if w <= frame_w and h <= frame_h:
                x1 = random.randint(0, frame_w - w)
                y1 = random.randint(0, frame_h - h)
                view = ViewPlan(frames, 0, (x1, y1, x1 + w, y1 + h), False, False)
                break
        if view is None:                                       # fallback: scale to `size`, centre crop
            resize, box = self._scale_center(frame_w, frame_h, self.size, self.size)
            view = ViewPlan(frames, 0, box, False, False, resize=resize)
        if random.random() < 0.3:                              # ClipColorJitter(p=0.3)
            ops = [("brightness", random.uniform(1 - 0.4, 1 + 0.4)), ("contrast", random.uniform(1 - 0.4, 1 + 0.4)),
                   ("saturation", random.uniform(1 - 0.4, 1 + 0.4)), ("hue", random.uniform(0 - 0.1, 0 + 0.1))]
            random.shuffle(ops)
            view.base, view.angle, view.jitter = True, 0.0, ops
        return ClipPlan(view)

    def plan_val(self, total_frames: int, frame_w: int, frame_h: int) -> ClipPlan:
        if self.short is None:
            raise ValueError("img_val is defined for sample_size 112 or 224 (preprocess_data.py:1140-1143)")
        frames = self._frames(total_frames)
        resize, box = self._scale_center(frame_w, frame_h, self.short, self.size)
        return ClipPlan(ViewPlan(frames, 0, box, False, False, resize=resize))

    def plan_test(self, total_frames: int, frame_w: int, frame_h: int) -> List[ClipPlan]:
        if self.short is None:
            raise ValueError("img_test is defined for sample_size 112 or 224")
        clip_range = (self.T - 1) * self.rate
        if total_frames - clip_range <= 0:
            seq, f = [], 1
            while len(seq) < self.T:
                seq.append(f)
                f += self.rate
                if f >= total_frames:
                    f = 1
            windows = [seq]
        else:
            windows = [[s_ + i * self.rate for i in range(self.T)] for s_ in range(1, total_frames - clip_range + 1, clip_range)]
            windows.append(list(range(total_frames - clip_range, total_frames + 1, self.rate)))
        resize, box = self._scale_center(frame_w, frame_h, self.short, self.size)
        return [ClipPlan(ViewPlan(list(w_), 0, box, False, False, resize=resize)) for w_ in windows]
