"""Numpy execution of `cstp_clip_view` descriptors: a CPU stand-in for csrc/clip_pipeline.cu with the same integer
arithmetic, used to check the HOST side of the clip pipeline (descriptor compilation, tap tables, fixed-point
coefficients) against the Pillow oracle without a GPU.  Test infrastructure only."""
import numpy as np

from cstp_b200 import lib as L

KS = 2 + L.CSTP_CLIP_KMAX
M32 = (1 << 32) - 1


def _clip8(v):
    return np.clip(v >> 22, 0, 255)


def _fetch_rotated_crop(frame, rot, box):
    """uint8 [ch][cw][3]: Image.transpose(ROTATE_*) then Image.crop(box) (outside = 0), as the kernel's index map."""
    H, W, _ = frame.shape
    x0, y0, x1, y1 = box
    Wr, Hr = (H, W) if rot & 1 else (W, H)
    ry, rx = np.mgrid[y0:y1, x0:x1]
    ok = (rx >= 0) & (rx < Wr) & (ry >= 0) & (ry < Hr)
    rxc, ryc = np.clip(rx, 0, Wr - 1), np.clip(ry, 0, Hr - 1)
    if rot == 0:
        xi, yi = rxc, ryc
    elif rot == 1:
        xi, yi = W - 1 - ryc, rxc
    elif rot == 2:
        xi, yi = W - 1 - rxc, H - 1 - ryc
    else:
        xi, yi = ryc, H - 1 - rxc
    out = frame[yi, xi]
    out[~ok] = 0
    return out


def _resample(img, tab, axis):
    a = img.astype(np.int64)
    if axis == 1:
        a = a.transpose(1, 0, 2)
    S = tab.shape[0]
    out = np.zeros((S,) + a.shape[1:], np.int64)
    for i in range(S):
        acc = np.full(a.shape[1:], 1 << 21, np.int64)
        for j in range(tab[i, 1]):
            acc += a[tab[i, 0] + j] * int(tab[i, 2 + j])
        out[i] = _clip8(acc)
    if axis == 1:
        out = out.transpose(1, 0, 2)
    return out.astype(np.uint8)


def _blend(deg, img, alpha):
    a = np.float32(alpha)
    t = deg.astype(np.float32) + a * (img.astype(np.float32) - deg.astype(np.float32))
    if 0.0 <= a <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def _luma(img):
    a = img.astype(np.int64)
    return ((a[..., 0] * 19595 + a[..., 1] * 38470 + a[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)


def _rgb2hsv(img):
    r, g, b = [img[..., i].astype(np.int32) for i in range(3)]
    maxc, minc = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    f32, f64 = np.float32, np.float64
    with np.errstate(divide="ignore", invalid="ignore"):
        cr = (maxc - minc).astype(f32)
        s = cr / maxc.astype(f32)
        rc, gc, bc = [(maxc - c).astype(f32) / cr for c in (r, g, b)]
        h = np.where(r == maxc, bc - gc, np.where(g == maxc, (2.0 + rc.astype(f64) - bc).astype(f32),
                                                  (4.0 + gc.astype(f64) - rc).astype(f32))).astype(f64)
        h = np.fmod(h / 6.0 + 1.0, 1.0).astype(f32)
        uh = np.clip((h.astype(f64) * 255.0).astype(np.int64), 0, 255)
        us = np.clip((s.astype(f64) * 255.0).astype(np.int64), 0, 255)
    gray = minc == maxc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def _hsv2rgb(hsv):
    f32, f64 = np.float32, np.float64
    h, s, v = [hsv[..., i].astype(f64) for i in range(3)]
    hf = h * 6.0 / 255.0
    i = np.floor(hf).astype(np.int64)
    f = (hf - i).astype(f32)
    fs = (s / 255.0).astype(f32)

    def rnd(x):
        return np.floor(x + 0.5)
    p = rnd(v * (1.0 - fs.astype(f64)))
    q = rnd(v * (1.0 - (fs * f).astype(f64)))
    t = rnd(v * (1.0 - fs.astype(f64) * (1.0 - f.astype(f64))))
    p, q, t = [np.clip(z, 0, 255).astype(np.uint8) for z in (p, q, t)]
    vv = hsv[..., 2]
    i6 = i % 6
    r = np.choose(i6, [vv, q, p, p, t, vv])
    g = np.choose(i6, [t, vv, vv, q, p, p])
    b = np.choose(i6, [p, p, t, vv, vv, q])
    gray = hsv[..., 1] == 0
    return np.stack([np.where(gray, vv, r), np.where(gray, vv, g), np.where(gray, vv, b)], -1).astype(np.uint8)


def _box_line(line, radius, ea, eb, ww, fw):
    n = line.shape[0]
    lastx = n - 1
    out = np.zeros_like(line)
    acc = (line[0] * (radius + 1)) & M32
    for x in range(ea - 1):
        acc = (acc + line[x]) & M32
    acc = (acc + line[lastx] * (radius - ea + 1)) & M32

    def step(x, sub, add, left, right):
        nonlocal acc
        acc = (acc + line[add] - line[sub]) & M32
        bulk = (acc * ww + (line[left] + line[right]) * fw) & M32
        out[x] = ((bulk + (1 << 23)) & M32) >> 24
    if ea <= eb:
        for x in range(0, ea):
            step(x, 0, x + radius, 0, x + radius + 1)
        for x in range(ea, eb):
            step(x, x - radius - 1, x + radius, x - radius - 1, x + radius + 1)
        for x in range(eb, lastx + 1):
            step(x, x - radius - 1, lastx, x - radius - 1, lastx)
    else:
        for x in range(0, eb):
            step(x, 0, x + radius, 0, x + radius + 1)
        for x in range(eb, ea):
            step(x, 0, lastx, 0, lastx)
        for x in range(ea, lastx + 1):
            step(x, x - radius - 1, lastx, x - radius - 1, lastx)
    return out


def _hblur(img, d):
    h, w, _ = img.shape
    line = img.transpose(1, 0, 2).reshape(w, -1)
    out = _box_line(line, d.blur_radius, d.blur_edge_a, d.blur_edge_b, int(d.blur_ww), int(d.blur_fw))
    return out.reshape(w, h, -1).transpose(1, 0, 2)


def run_view(d: "L.ClipView", coef: np.ndarray, video: np.ndarray, T: int, S: int) -> np.ndarray:
    """Executes one descriptor on a host copy of the video; returns the (3, T, S, S) fp32 clip."""
    out = np.zeros((3, T, S, S), np.float32)
    for t in range(T):
        img = _fetch_rotated_crop(video[d.frames[t]], d.rot, tuple(d.box))
        img = _resample(img, coef[0], 1)
        img = _resample(img, coef[1], 0)
        if d.rotate:
            a0, a1, a2, a3, a4, a5 = list(d.rot_fix)
            y, x = np.mgrid[0:S, 0:S]
            xin, yin = (a2 + x * a0 + y * a1) >> 16, (a5 + x * a3 + y * a4) >> 16
            ok = (xin >= 0) & (xin < S) & (yin >= 0) & (yin < S)
            r = np.zeros_like(img)
            r[ok] = img[yin[ok], xin[ok]]
            img = r
        for j in range(d.n_jitter):
            op, f = d.jitter_op[j], d.jitter_f[j]
            if op == 0:
                img = _blend(np.zeros_like(img), img, f)
            elif op == 1:
                lum = _luma(img)
                mean = int(float(lum.astype(np.int64).sum()) / lum.size + 0.5)
                img = _blend(np.full_like(img, mean), img, f)
            elif op == 2:
                lum = _luma(img)
                img = _blend(np.stack([lum, lum, lum], -1), img, f)
            else:
                hsv = _rgb2hsv(img)
                hsv[..., 0] = ((hsv[..., 0].astype(np.int32) + d.hue_shift) & 255).astype(np.uint8)
                img = _hsv2rgb(hsv)
        if d.gray[t] >= 0:
            c = img[..., d.gray[t]]
            img = np.stack([c, c, c], -1)
        if d.blur:
            a = img.astype(np.int64)
            for _ in range(3):
                a = _hblur(a, d)
            a = a.transpose(1, 0, 2)
            for _ in range(3):
                a = _hblur(a, d)
            img = a.transpose(1, 0, 2).astype(np.uint8)
        if d.flip:
            img = img[:, ::-1]
        f = img.astype(np.float32) / np.float32(255)
        out[:, t] = (f * np.float32(2) - np.float32(1)).transpose(2, 0, 1)
    return out
