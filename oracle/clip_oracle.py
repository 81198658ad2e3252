"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's pretraining clip pipeline for one sample.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path
(cstp_b200/data_process) never does.

What it restates (reference file:line):
  * data_process/datasets.py:859-948   UcfRepreBYOLSpPre.repre_train_clip   -- which frames, which 90-degree rotation
  * data_process/datasets.py:1308-1405 Kin400RepreLMDB.repre_train_clip      -- the 0-based variant (clip 2 re-reads clip 1's
                                                                               frames: `raw[start_frame + i]` at :1397)
  * data_process/preprocess_data.py:479-565  ClipRandomSizedCropOverlap      -- crop boxes, bicubic resize to 112
  * data_process/preprocess_data.py:1103-1130 get_transforms('pre_train')    -- null / base transform chains
  * data_process/datasets.py:951-1097 UcfFineTune + preprocess_data.py:1131-1155 get_transforms('img' | 'img_val' |
    'img_test'), :440-476 ClipRandomSizedCrop, :843-866 ClipScale, :815-840 ClipCenterCrop -- finetune / val / test clips
The pixel arithmetic of the reference lives in third-party Pillow / torchvision (unpinned by the reference; this image:
Pillow 12.2.0, torchvision 0.26.0).  `render_plan` therefore replays a *plan* (every random decision already taken,
see cstp_b200/data_process/clip_plan.py) through the same Pillow / torchvision calls in the same order; it is pinned by
tests/test_clip_pipeline.py against clips produced by the unmodified reference classes (oracle/make_golden_clips.py ->
tests/golden/clips_ref.npz; oracle/make_golden_clips_ft.py -> clips_ft_ref.npz): bit-exact.
"""
from __future__ import annotations

import numpy as np
import torch
from PIL import Image, ImageFilter

ROT_METHOD = (None, Image.ROTATE_90, Image.ROTATE_180, Image.ROTATE_270)      # datasets.py:19  ROTATE = [0, 2, 3, 4]


from cstp_b200.synthetic import synthetic_video  # noqa: E402,F401  (seeded test video; no algorithm in it)


def _to_tensor_tf(img: Image.Image) -> torch.Tensor:
    """transforms.ToTensor + Normalize(flag='tf') -- preprocess_data.py:434-437, 353-362."""
    t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div(255)
    return torch.clamp(t * 2.0 - 1.0, -1.0, 1.0)


def render_view(view, video: np.ndarray, frame_base: int, size: int = 112) -> torch.Tensor:
    """One view of a plan -> (3, T, size, size) fp32.  `video[n - frame_base]` is the frame the reference opens as number n."""
    import torchvision.transforms.functional as F
    imgs = []
    for n in view.frames:
        img = Image.fromarray(video[n - frame_base], "RGB")
        if ROT_METHOD[view.rot] is not None:
            img = img.transpose(ROT_METHOD[view.rot])                                # datasets.py:896-903, 931-932
        imgs.append(img)
    x0, y0, x1, y1 = view.box
    if getattr(view, "resize", None) is not None:       # ClipScale + ClipCenterCrop (preprocess_data.py:843-866, 829-840)
        ow, oh = view.resize
        imgs = [(i if i.size == (ow, oh) else i.resize((ow, oh), Image.BICUBIC)).crop((x0, y0, x1, y1)) for i in imgs]
    else:
        imgs = [i.crop((x0, y0, x1, y1)).resize((size, size), Image.BICUBIC) for i in imgs]   # preprocess_data.py:514-515
    if view.base:                                                                   # preprocess_data.py:1112-1122
        if view.angle:                               # RandomRotation :1095 (the finetune chain has none: angle 0.0)
            imgs = [i.rotate(view.angle) for i in imgs]
        if view.jitter is not None:                                                  # ClipColorJitter :659-663
            fn = dict(brightness=F.adjust_brightness, contrast=F.adjust_contrast, saturation=F.adjust_saturation,
                      hue=F.adjust_hue)
            for name, factor in view.jitter:
                imgs = [fn[name](i, factor) for i in imgs]
        if view.gray is not None:                                                    # ClipRandomGray :705-711
            out = []
            for i, ch in zip(imgs, view.gray):
                a = np.array(i)[:, :, ch]
                out.append(Image.fromarray(np.dstack([a, a, a]), "RGB"))
            imgs = out
        if view.blur_sigma is not None:                                              # ClipGaussianBlur :681-688
            imgs = [i.filter(ImageFilter.GaussianBlur(radius=view.blur_sigma)) for i in imgs]
    if view.flip:                                                                    # ClipRandomHorizontalFlip :577-582
        imgs = [i.transpose(Image.FLIP_LEFT_RIGHT) for i in imgs]
    return torch.stack([_to_tensor_tf(i) for i in imgs]).transpose(0, 1).contiguous()


def render_plan(plan, video: np.ndarray, size: int = 112):
    """SamplePlan -> (clip_1, clip_2), each (3, T, size, size) fp32, as `__getitem__` returns them (datasets.py:850-857)."""
    return tuple(render_view(v, video, plan.frame_base, size) for v in plan.views)
